#!/bin/bash
# Round 2, GPU call T (1 GPU): per-launch times of an LSQR and a SYMMLQ trip (ncu launch list).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_lls_launches.csv python scripts/prof_lls.py > gpurun_out/r2t_prof.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2t_prof.log
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(open("gpurun_out/r2t_lls_launches.csv", errors="ignore")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    per.setdefault(r[ii], {"k": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
for i, d in list(per.items())[-60:]:
    print(i, "%8.1f us  rd %7.1f MB  wr %7.1f MB  %s" % (d.get("gpu__time_duration.sum", 0) / 1e3, d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6, re.sub(r"\(.*", "", d["k"])[:90]))
PY
