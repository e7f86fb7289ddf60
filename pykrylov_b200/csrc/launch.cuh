// launch.cuh -- host-side launch helpers shared by ops.cu and the solvers.
#pragma once

#include <stdlib.h>

#include "spmv.cuh"

constexpr int KRY_DEFAULT_TILE    = 4096;
constexpr int KRY_DEFAULT_THREADS = 256;
constexpr int KRY_TMA_STAGES      = 3;
constexpr int KRY_HALO_DEDICATED_CTAS = 0;   // default of KRY_HALO_PUSH_CTAS (measured, profiles/r2p_*)

int csr_build_partition(kry_ctx *c, CsrDev &m, int tile_nnz);
int kry_halo_exchange(kry_csr *M, double *x_dev);    // comm.cu; no-op unless sharded
int kry_halo_exchange_dir(kry_csr *M, double *x_dev, const double *r_dev, const double *beta_dev);

template <class K>
static inline int set_max_smem(K kernel, kry_ctx *c, size_t bytes)
{
    (void)c;
    if (bytes <= 48 * 1024) return KRY_OK;
    // the opt-in limit counts static + dynamic smem, so ask for exactly what is used
    KRY_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return KRY_OK;
}

// Persistent grid-stride kernels run exactly one resident wave: SMs x the occupancy
// the kernel really gets (queried once per instantiation), never more.
template <class K>
static inline int vec_grid(kry_ctx *c, int64_t n, K kernel)
{
    static int per_sm = 0;
    if (per_sm == 0) {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, 256, 0) != cudaSuccess || b < 1) b = 4;
        per_sm = b;
    }
    int64_t need = (n + 255) / 256;
    int64_t cap = (int64_t)c->sm_count * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// y-side work is described by Epi, x-side by Gather; ND fused inner products are
// reduced on device and handed to Fin (see common.cuh).
template <int ND, class Gather, class Epi, class Fin>
int spmv_launch(kry_csr *M, bool trans, Gather g, Epi epi, Fin fin, const int *done, int defer)
{
    kry_ctx *c = M->ctx;
    CsrDev *m = &M->A;
    if (trans && !(M->flags & KRY_CSR_SYMMETRIC)) {
        KRY_REQUIRE(M->has_T, KRY_ERR_STATE, "spmv: transpose requested but not built "
                    "(create with KRY_CSR_BUILD_TRANSPOSE or call kry_csr_build_transpose)");
        m = &M->T;
    }
    // AUTO: measured on B200 (profiles/): for short rows one thread per row with direct,
    // L1-cached loads beats both smem-staged variants (5.4 vs 3.8 TB/s on the 5-pt
    // Laplacian); rows long enough to serialise a thread go to the nnz-stream kernel.
    int kind = M->kind;
    if (kind == KRY_SPMV_AUTO) kind = (m->max_row <= 64) ? KRY_SPMV_ROW : KRY_SPMV_STREAM;
#ifdef KRY_EMULATE
    if (kind == KRY_SPMV_STREAM || kind == KRY_SPMV_TMA || kind == KRY_SPMV_ROWPF2) kind = KRY_SPMV_ROW;   // host emulation: row loops only
#endif
    // The TMA-staged row kernel is an experiment for the stand-alone products (kry_spmv / kry_spmv_dot):
    // bit-exact there, but no faster than the plain row kernel (0.1317 vs 0.1292 ms on config 2) and, inside
    // the graph-replayed CG loop with exactly one resident wave, it showed multi-millisecond stalls that are
    // not understood (profiles/r2k_tune.log).  The solver loops therefore never use it.
    if (kind == KRY_SPMV_ROWPF2 && done != c->never_done && done != c->gate) kind = KRY_SPMV_ROW;
    const int tile = M->tile_nnz ? M->tile_nnz : KRY_DEFAULT_TILE;
    const int threads = M->threads ? M->threads : KRY_DEFAULT_THREADS;
    const size_t budget = (size_t)c->smem_optin - 2048;   // static smem of the reduction + barriers
    size_t smem = 0;
    int grid = 1, cap = 0;

    if (kind == KRY_SPMV_STREAM) {
        smem = ((size_t)tile + m->max_row) * sizeof(double);
        if (smem > budget) kind = KRY_SPMV_ROW;
    } else if (kind == KRY_SPMV_TMA) {
        cap = (tile + m->max_row + 8 + 3) & ~3;
        smem = (size_t)KRY_TMA_STAGES * cap * 12;
        if (smem > budget) kind = KRY_SPMV_ROW;
    }
    if (kind == KRY_SPMV_ROW || kind == KRY_SPMV_ROWB8 || kind == KRY_SPMV_ROWB4 || kind == KRY_SPMV_ROWPF ||
        kind == KRY_SPMV_ROWPF2) {
        int64_t need = (m->nrows + 255) / 256;
        if (need < 1) need = 1;
        // one resident wave (8 CTAs x 256 threads per SM) measured best: 6.25 vs 5.5 TB/s with
        // 64 CTAs/SM on the 5-pt Laplacian; `threads`/32 overrides the CTAs per SM for sweeps
        int per_sm = (M->threads >= 64 && M->tile_nnz == 0) ? M->threads / 32 : 8;
        if ((kind == KRY_SPMV_ROW || kind == KRY_SPMV_ROWPF || kind == KRY_SPMV_ROWPF2) && per_sm == 8) {
            // ... but never more CTAs than are really resident (an epilogue above 32 registers
            // would otherwise leave a second, partial wave)
            static int occs[3] = {0, 0, 0};               // per instantiation: row, rowpf, rowpf2
            int &occ = occs[kind == KRY_SPMV_ROW ? 0 : (kind == KRY_SPMV_ROWPF ? 1 : 2)];
            if (occ == 0) {
                int b = 0;
                cudaError_t qe;
                if (kind == KRY_SPMV_ROW)
                    qe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, spmv_row_kernel<ND, Gather, Epi, Fin>, 256, 0);
                else if (kind == KRY_SPMV_ROWPF)
                    qe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, spmv_rowpf_kernel<ND, 1, Gather, Epi, Fin>, 256, 0);
                else
#ifdef KRY_EMULATE
                    qe = cudaErrorUnknown;
#else
                    qe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, spmv_rowtma_kernel<ND, Gather, Epi, Fin>, 256, 0);
#endif
                if (qe != cudaSuccess || b < 1) b = 8;
                occ = b;
            }
            if (occ < per_sm) per_sm = occ;
        }
        const int64_t capg = (int64_t)c->sm_count * per_sm;
        grid = (int)(need < capg ? need : capg);
    } else {
        KRY_TRY(csr_build_partition(c, *m, tile));
        grid = m->nblocks;
        if (kind == KRY_SPMV_TMA) {
            int per_sm = (int)((size_t)(c->smem_optin) / (smem + 1024));
            if (per_sm > 2048 / threads) per_sm = 2048 / threads;
            if (per_sm < 1) per_sm = 1;
            const int g2 = c->sm_count * per_sm;
            if (grid > g2) grid = g2;
        }
    }
    KRY_TRY(kry_ctx_ensure_partials(c, grid));
    ReduceWs ws = kry_ws(c);
    ws.defer = (defer == 1);
    ws.p2p = (defer == 2);
    CsrView A = csr_view(*m);
    A.hints = (c->l2_hints & 2) ? 1 : 0;     // bit 1: evict_first on the CSR streams (measured harmful)
    const bool prof = ND > 0 && c->prof_ev && c->prof_n < c->prof_cap;
    if (prof) KRY_CUDA(cudaEventRecord(c->prof_ev[2 * c->prof_n], c->stream));

#ifdef KRY_EMULATE
    (void)cap;
    (void)smem;
    (void)threads;
    if (kind == KRY_SPMV_ROW)
        emu_launch<ND>(grid, 256, ws, fin, [&] { spmv_row_kernel<ND, Gather, Epi, Fin>(A, g, epi, ws, fin, done); });
    else if (kind == KRY_SPMV_ROWPF)
        emu_launch<ND>(grid, 256, ws, fin, [&] { spmv_rowpf_kernel<ND, 1, Gather, Epi, Fin>(A, g, epi, ws, fin, done); });
    else if (kind == KRY_SPMV_ROWB8)
        emu_launch<ND>(grid, 256, ws, fin, [&] { spmv_rowb_kernel<ND, 8, Gather, Epi, Fin>(A, g, epi, ws, fin, done); });
    else
        emu_launch<ND>(grid, 256, ws, fin, [&] { spmv_rowb_kernel<ND, 4, Gather, Epi, Fin>(A, g, epi, ws, fin, done); });
#else
    if (kind == KRY_SPMV_ROW) {
        spmv_row_kernel<ND, Gather, Epi, Fin><<<grid, 256, 0, c->stream>>>(A, g, epi, ws, fin, done);
    } else if (kind == KRY_SPMV_ROWPF) {
        spmv_rowpf_kernel<ND, 1, Gather, Epi, Fin><<<grid, 256, 0, c->stream>>>(A, g, epi, ws, fin, done);
    } else if (kind == KRY_SPMV_ROWPF2) {        // TMA-staged row pointers + bulk L2 prefetch of the next tile's CSR window
        spmv_rowtma_kernel<ND, Gather, Epi, Fin><<<grid, 256, 0, c->stream>>>(A, g, epi, ws, fin, done);
    } else if (kind == KRY_SPMV_ROWB8) {
        spmv_rowb_kernel<ND, 8, Gather, Epi, Fin><<<grid, 256, 0, c->stream>>>(A, g, epi, ws, fin, done);
    } else if (kind == KRY_SPMV_ROWB4) {
        spmv_rowb_kernel<ND, 4, Gather, Epi, Fin><<<grid, 256, 0, c->stream>>>(A, g, epi, ws, fin, done);
    } else if (kind == KRY_SPMV_STREAM) {
        auto k = spmv_stream_kernel<ND, Gather, Epi, Fin>;
        KRY_TRY(set_max_smem(k, c, smem));
        k<<<grid, threads, smem, c->stream>>>(A, g, epi, ws, fin, done);
    } else {
        auto k = spmv_tma_kernel<ND, KRY_TMA_STAGES, Gather, Epi, Fin>;
        KRY_TRY(set_max_smem(k, c, smem));
        k<<<grid, threads, smem, c->stream>>>(A, cap, g, epi, ws, fin, done);
    }
#endif
    if (prof) {
        KRY_CUDA(cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], c->stream));
        c->prof_n++;
    }
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}

// Row shards, KRY_OPT_HALO_P2P: the halo exchange rides in the SpMV launch (spmv.cuh,
// spmv_row_shard_kernel).  `tbl` is the device table of the gathered vector (solver.cuh); the
// fused inner products are all-reduced in-kernel over the same peer mapping.  Every rank issues
// the same sequence of these launches, so the host-side tag counter agrees everywhere.
static inline bool spmv_shard_row_kind(const kry_csr *M)
{
    int kind = M->kind;
    if (kind == KRY_SPMV_AUTO) kind = (M->A.max_row <= 64) ? KRY_SPMV_ROW : KRY_SPMV_STREAM;
    return kind == KRY_SPMV_ROW && M->halo.active;
}
static inline bool spmv_shard_fusable(const kry_csr *M)
{
    return spmv_shard_row_kind(M) && M->ctx->p2p_on && M->ctx->halo_p2p;
}

// tbl != nullptr: exchange fused into the launch (needs the in-kernel all-reduce, defer = 2);
// tbl == nullptr: the halo tail was filled before the launch (pack kernel + ncclAllGather) -- same
// kernel, same row order, push and wait switched off.  defer as in spmv_launch.
template <int ND, class Gather, class Epi, class Fin>
int spmv_shard_launch(kry_csr *M, Gather g, Epi epi, Fin fin, const int *done, const HaloTable *tbl, int defer)
{
    static_assert(ND > 0, "the in-kernel all-reduce at the end of the launch is what orders the exchanges");
    kry_ctx *c = M->ctx;
    CsrDev *m = &M->A;
    const HaloPlan &hp = M->halo;
    int64_t need = (m->nrows + 255) / 256;
    if (need < 1) need = 1;
    static int occ = 0;              // per instantiation
    if (occ == 0) {
        int b = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, spmv_row_shard_kernel<ND, Gather, Epi, Fin>, 256, 0) !=
                cudaSuccess || b < 1)
            b = 8;
        occ = b < 8 ? b : 8;
    }
    // a few CTAs of the wave only publish the boundary entries (same grid whichever way the halo travels,
    // so the fused inner products are the same bits); one resident wave: publishing and waiting CTAs
    // must be co-resident
    // KRY_HALO_PUSH_CTAS (A/B switch, read once): N > 0 = N dedicated publishing CTAs that own no rows;
    // 0 = the last ceil(n_send/256) row CTAs publish before they walk their rows
    static int dedicated = -1;
    if (dedicated < 0) {
        const char *e = getenv("KRY_HALO_PUSH_CTAS");
        dedicated = e ? atoi(e) : KRY_HALO_DEDICATED_CTAS;
    }
    int push = (hp.n_send + 255) / 256;
    int64_t capg = (int64_t)c->sm_count * occ;
    if (dedicated > 0) {
        if (push > dedicated) push = dedicated;
        capg -= push;
    }
    if (capg < 1) capg = 1;
    const int row_ctas = (int)(need < capg ? need : capg);
    const int grid = dedicated > 0 ? row_ctas + push : row_ctas;
    if (push > grid) push = grid;
    KRY_TRY(kry_ctx_ensure_partials(c, grid));
    ReduceWs ws = kry_ws(c);
    ws.defer = (defer == 1);
    ws.p2p = (defer == 2);
    CsrView A = csr_view(*m);
    HaloArgs h;
    h.tbl = tbl;
    h.send_idx = hp.send_idx;
    h.n_send = hp.n_send;
    h.push_ctas = push;
    h.row_ctas = row_ctas;
    h.ticket = c->counter + 24;                             // inside the zeroed 256-byte counter block
    h.tag = tbl ? ++c->halo_seq : 0ull;
    h.rot = hp.rot;
    h.v_wait = hp.v_wait;
    h.skip_push = tbl ? 0 : 3;
    h.emu_phase = 0;
    h.trace = tbl ? c->halo_trace : nullptr;
    KRY_REQUIRE(!tbl || defer == 2, KRY_ERR_STATE, "fused halo exchange without the in-kernel all-reduce");
    const bool prof = c->prof_ev && c->prof_n < c->prof_cap;
    if (prof) KRY_CUDA(cudaEventRecord(c->prof_ev[2 * c->prof_n], c->stream));
#ifdef KRY_EMULATE
    if (emu_fibers_on) {
        // SIMT mode: all blocks alive at once (a block that waits for a peer's flag must not keep
        // this rank's later push blocks from running)
        auto body = [&] { spmv_row_shard_kernel<ND, Gather, Epi, Fin>(A, g, epi, ws, fin, done, h); };
        auto call = [](const void *k) { (*static_cast<const decltype(body) *>(k))(); };
        h.emu_phase = 1;                                   // the publishing blocks first (they wait for nobody) ...
        emu_launch_fibers_mode(grid, 256, &body, call, 0);
        h.emu_phase = 2;                                   // ... then the row blocks
        emu_launch_fibers_mode(grid, 256, &body, call, 1);
    } else {
        // fast mode plays the threads one after the other: the push (all of it, then the flags) is
        // played first by the launcher with the kernel's own device function
        if (!*done && tbl) {
            Gather g0 = g;
            g0.init();
            gridDim = EmuDim{(unsigned)grid, 1, 1};
            blockDim = EmuDim{256, 1, 1};
            emu_halo_push_all(h, g0);
        }
        h.skip_push |= 1;
        emu_launch<ND>(grid, 256, ws, fin, [&] { spmv_row_shard_kernel<ND, Gather, Epi, Fin>(A, g, epi, ws, fin, done, h); });
    }
#else
    spmv_row_shard_kernel<ND, Gather, Epi, Fin><<<grid, 256, 0, c->stream>>>(A, g, epi, ws, fin, done, h);
#endif
    if (prof) {
        KRY_CUDA(cudaEventRecord(c->prof_ev[2 * c->prof_n + 1], c->stream));
        c->prof_n++;
    }
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}

template <int ND, class Body, class Fin>
int vec_pass_launch(kry_ctx *c, int64_t n, Body body, Fin fin, const int *done, int defer)
{
    const int grid = vec_grid(c, n, vec_pass_kernel<ND, Body, Fin>);
    KRY_TRY(kry_ctx_ensure_partials(c, grid));
    ReduceWs ws = kry_ws(c);
    ws.defer = (defer == 1);
    ws.p2p = (defer == 2);
#ifdef KRY_EMULATE
    emu_launch<ND>(grid, 256, ws, fin, [&] { vec_pass_kernel<ND, Body, Fin>(n, body, ws, fin, done); });
#else
    vec_pass_kernel<ND, Body, Fin><<<grid, 256, 0, c->stream>>>(n, body, ws, fin, done);
#endif
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}

template <class Body>
int vec_map_launch(kry_ctx *c, int64_t n, Body body, const int *done)
{
#ifdef KRY_EMULATE
    emu_launch<0>(vec_grid(c, n, vec_map_kernel<Body>), 256, ReduceWs(), NoFin(), [&] { vec_map_kernel<Body>(n, body, done); });
#else
    vec_map_kernel<Body><<<vec_grid(c, n, vec_map_kernel<Body>), 256, 0, c->stream>>>(n, body, done);
#endif
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}

template <class Fin>
int finalize_launch(kry_ctx *c, Fin fin, const int *done)
{
#ifdef KRY_EMULATE
    emu_launch<0>(1, 32, ReduceWs(), NoFin(), [&] { finalize_kernel<Fin>(fin, c->sums, done); });
#else
    finalize_kernel<Fin><<<1, 32, 0, c->stream>>>(fin, c->sums, done);
#endif
    c->launches++;
    KRY_CUDA(cudaGetLastError());
    return KRY_OK;
}
