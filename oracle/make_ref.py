#!/usr/bin/env python3
"""Generate oracle/_ref/refpykrylov: a runnable Python-3 image of the reference.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is ever imported by the
product package (pykrylov_b200); only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may use it.

The reference (/root/reference, PythonOptimizers/pykrylov v0.2.1) is Python-2
only and cannot be imported by the Python 3.12 interpreter of this image.  This
script reads the reference sources *where they lie* and writes a mechanically
transliterated copy to the git-ignored directory oracle/_ref/ (it travels to
the GPU box with the gpurun snapshot, it never enters history).  No algorithmic
line is changed; the edits are exactly the syntax-level ones listed in
SURVEY.md Appendix A:

  1. implicit relative imports      -> explicit (``from cg import *`` -> ``from .cg import *``)
  2. ``__all__ = filter(...)``      -> ``list(filter(...))``; ``__all__ += 'x'`` -> ``+= ['x']``
  3. removed NumPy aliases          -> np.float -> np.float64 etc.
  4. ``xrange``                     -> ``range``
  5. ``print x``                    -> ``print(x)``
  6. ``map``/``reduce`` py3 forms in blkop.py
  7. ``__main__`` demo blocks dropped
  8. the package is renamed ``refpykrylov`` so that it can coexist in one
     interpreter with the product's ``pykrylov`` drop-in alias
  9. ONE bug fix (not syntax): symmlq.py ``self.matvec(v)`` -> ``self.op * v``
      -- without it Symmlq raises AttributeError.

Usage:  python oracle/make_ref.py [--src /root/reference] [--dst oracle/_ref]
"""
import argparse
import os
import re
import shutil
import sys

PKG_OLD = "pykrylov"
PKG_NEW = "refpykrylov"


def _split_comment(code):
    """Split ``code`` into (statement, trailing comment) outside of strings."""
    quote = None
    i = 0
    while i < len(code):
        c = code[i]
        if quote:
            if c == "\\":
                i += 2
                continue
            if c == quote:
                quote = None
        elif c in "'\"":
            quote = c
        elif c == "#":
            return code[:i].rstrip(), "  " + code[i:]
        i += 1
    return code.rstrip(), ""


def _split_semicolons(code):
    """Split a physical line on top-level ``;`` (outside strings / brackets)."""
    parts, cur, quote, depth = [], [], None, 0
    i = 0
    while i < len(code):
        c = code[i]
        if quote:
            cur.append(c)
            if c == "\\" and i + 1 < len(code):
                cur.append(code[i + 1])
                i += 2
                continue
            if c == quote:
                quote = None
        elif c in "'\"":
            quote = c
            cur.append(c)
        elif c in "([{":
            depth += 1
            cur.append(c)
        elif c in ")]}":
            depth -= 1
            cur.append(c)
        elif c == ";" and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(c)
        i += 1
    tail = "".join(cur).strip()
    if tail:
        parts.append(tail)
    return parts


def _py3_print(stmt):
    """``print a, b`` -> ``print(a, b)``; bare ``print`` -> ``print()``."""
    if stmt == "print":
        return "print()"
    if stmt.startswith("print ") or stmt.startswith("print\t"):
        body = stmt[5:].strip()
        if body.endswith(","):                     # py2 "no newline" form
            return "print(%s end=' ')" % body
        return "print(%s)" % body
    return stmt


def convert_prints(lines):
    out = []
    i = 0
    while i < len(lines):
        line = lines[i]
        stripped = line.lstrip()
        indent = line[: len(line) - len(stripped)]
        inline = re.match(r"((?:if|elif|else|for|while)\b[^#'\"]*?:\s+)(print(?:\s.*|$))", stripped)
        if inline:
            # one-line compound statement:  ``if cond: print x``
            stmt, comment = _split_comment(inline.group(2).rstrip("\n"))
            pieces = [_py3_print(p) for p in _split_semicolons(stmt)]
            out.append(indent + inline.group(1) + "; ".join(pieces) + comment + "\n")
        elif re.match(r"print(\s|$)", stripped):
            # gather backslash continuations into one logical statement
            logical = stripped.rstrip("\n")
            while logical.rstrip().endswith("\\") and i + 1 < len(lines):
                i += 1
                logical = logical.rstrip()[:-1].rstrip() + " " + lines[i].strip()
            stmt, comment = _split_comment(logical)
            pieces = [_py3_print(p) for p in _split_semicolons(stmt)]
            out.append(indent + "; ".join(pieces) + comment + "\n")
        else:
            out.append(line)
        i += 1
    return out


def drop_main_block(lines):
    for k, line in enumerate(lines):
        if re.match(r"if\s+__name__\s*==\s*['\"]__main__['\"]\s*:", line):
            return lines[:k]
    return lines


NUMPY_ALIASES = [
    (r"\bnp\.float_\b", "np.float64"),
    (r"\bnp\.float\b(?!\d)", "np.float64"),
    (r"\bnp\.complex_\b", "np.complex128"),
    (r"\bnp\.complex\b(?!\d)", "np.complex128"),
    (r"\bnp\.int\b(?!\w)", "np.int64"),
]


def convert_source(text, siblings, relpath):
    lines = text.splitlines(keepends=True)
    if not relpath.endswith(os.path.join("tests", "")):
        lines = drop_main_block(lines)
    lines = convert_prints(lines)
    text = "".join(lines)

    # 1. implicit relative imports of sibling modules / packages
    def rel_from(m):
        mod = m.group(2)
        head = mod.split(".")[0]
        if head in siblings:
            return "%sfrom .%s import" % (m.group(1), mod)
        return m.group(0)

    text = re.sub(r"(?m)^(\s*)from\s+([\w\.]+)\s+import", rel_from, text)

    # 8. package rename
    text = re.sub(r"\bfrom\s+%s(\.|\s)" % PKG_OLD, r"from %s\1" % PKG_NEW, text)
    text = re.sub(r"\bimport\s+%s\b" % PKG_OLD, "import %s" % PKG_NEW, text)

    # 2. __all__ idioms
    text = re.sub(r"__all__\s*=\s*filter\((.*)\)\s*$", r"__all__ = list(filter(\1))",
                  text, flags=re.M)
    text = text.replace("__all__ += '__version__'", "__all__ += ['__version__']")

    # 3. numpy aliases, 4. xrange
    for pat, rep in NUMPY_ALIASES:
        text = re.sub(pat, rep, text)
    text = re.sub(r"\bxrange\b", "range", text)

    # 6. blkop.py: map() is lazy / reduce moved in py3
    if relpath.endswith("blkop.py"):
        text = re.sub(r"(?<![\w\.])map\(([^\n]*)\)\s*$", r"list(map(\1))", text, flags=re.M)
        if "reduce(" in text and "from functools import reduce" not in text:
            text = "from functools import reduce\n" + text

    # 9. the one bug fix
    if relpath.endswith(os.path.join("symmlq", "symmlq.py")):
        text = text.replace("self.matvec(v)", "self.op * v")

    # tests only: deprecated unittest alias, float index arrays
    if os.sep + "tests" + os.sep in relpath:
        text = text.replace("self.assert_(", "self.assertTrue(")
    return text


def generate(src, dst):
    src_pkg = os.path.join(src, PKG_OLD)
    if not os.path.isdir(src_pkg):
        raise SystemExit("reference package not found at %s" % src_pkg)
    dst_pkg = os.path.join(dst, PKG_NEW)
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    n_files = 0
    for root, dirs, files in os.walk(src_pkg):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        rel = os.path.relpath(root, src_pkg)
        out_dir = os.path.normpath(os.path.join(dst_pkg, rel))
        os.makedirs(out_dir, exist_ok=True)
        siblings = {f[:-3] for f in files if f.endswith(".py")} | set(dirs)
        for f in files:
            if not f.endswith(".py") or f == "setup.py":
                continue
            with open(os.path.join(root, f), "r", encoding="utf-8", errors="replace") as fh:
                text = fh.read()
            relpath = os.path.join(rel, f)
            new = convert_source(text, siblings - {f[:-3]} if f != "__init__.py" else siblings,
                                 relpath)
            with open(os.path.join(out_dir, f), "w", encoding="utf-8") as fh:
                fh.write(new)
            n_files += 1
    with open(os.path.join(dst, "README.txt"), "w") as fh:
        fh.write("Generated by oracle/make_ref.py from %s -- do not edit, do not commit.\n" % src)
    return n_files


def main(argv=None):
    here = os.path.dirname(os.path.abspath(__file__))
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=os.path.join(here, "_ref"))
    args = ap.parse_args(argv)
    n = generate(args.src, args.dst)
    print("wrote %d files under %s" % (n, os.path.join(args.dst, PKG_NEW)))
    # sanity: it must import
    sys.path.insert(0, args.dst)
    import importlib
    mod = importlib.import_module(PKG_NEW)
    print("import ok:", mod.__name__, getattr(mod, "__version__", "?"))


if __name__ == "__main__":
    main()
