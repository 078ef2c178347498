// env.cuh -- the world model in HBM and its shared-memory staging.
//
// Layout in HBM (one blob per precision, built once by auvrrt_env_create):
//   EnvHeader | circles SoA (cx, cy, r, r_eff, r_eff^2) | polygon (px, py) | habitats SoA |
//   time bins (b0, b1) | cell index (breakpoints, piece offsets, candidates) | probs [T][C]
// Everything up to `hot_bytes` is copied into shared memory by ONE TMA bulk copy
// (cp.async.bulk.shared::cluster.global, completion on an mbarrier) at kernel start; the
// probability table follows it into shared memory too when it fits, else it is read from L2.
//
// Semantics preserved from the reference:
//  * r_eff[j] = max_{k >= j} r[k].  RRT.check_collision never resets dList between obstacles
//    (rrt_dubins.py:535-542), so it reports a collision iff  exists k: min_{j<=k} m_j <= r_k
//    <=> exists j: m_j <= max_{k>=j} r_k.  The running-minimum quirk is therefore EXACTLY an
//    inflation of circle j to the suffix maximum of the radii, in obstacle order.
//  * cell index: cost.py:181-184 scans the cell dict in order and takes the first cell with
//    x>=c0 and x<=c2 and y>=c1 and x<=c3 (sic).  The x-conditions say x in [c0, min(c2,c3)].
//    The breakpoints {c0, min(c2,c3)} cut the x axis into points and open intervals ("pieces") on
//    which the set of x-admissible cells is constant; within a piece, a cell can only be the
//    first match if its c1 is strictly below the c1 of every earlier admissible cell, so each
//    piece keeps that strictly decreasing candidate list.  First candidate with c1 <= y wins.
#pragma once
#include "common.cuh"

namespace auv {

struct EnvHeader {
    int K, E, H, T, C, NB, NP, NCAND;
    int convex;        // +1: convex CCW, -1: convex CW, 0: general simple polygon
    int hot_bytes;     // bytes [0, hot_bytes) are staged in shared memory
    int total_bytes;
    int off_cx, off_cy, off_cr, off_creff, off_creff2;
    int off_px, off_py;
    int off_hx, off_hy, off_hr, off_hr2;
    int off_b0, off_b1;
    int off_brk, off_piece, off_c1, off_cell;
    int off_probs;
    int off_xb;        // x-bucket table for the shark-cell pieces (u16 per bucket), staged
    int nxb;           // number of x buckets (0: none)
    int off_grid;      // classification grid (3 planes of u32, one word per cell each), not staged: read through L1/L2
    int gnx, gny;      // grid dimensions (0: no grid)
    int bins_uniform;  // 1: bins are [s0 + i w, s0 + (i+1) w], contiguous and in order
    int off_xcell;     // fine x table for the shark-cell lookup (XCell<R> per bucket), bytes [total_bytes, ext_bytes)
    int nxc;           // buckets in it (0: none)
    int ext_bytes;     // a kernel with shared memory to spare stages [0, ext_bytes)
    double bbox[4];    // minx, miny, maxx, maxy of the polygon (Polygon.bounds, rrt_dubins.py:334)
    double gx0, gy0, gs;   // grid origin and cell size
    double bin_s0, bin_w;
    double xb0, xbw;       // x-bucket origin and width
    double xc0, xcw;       // fine x table origin and bucket width
};
static_assert(sizeof(EnvHeader) % 16 == 0, "EnvHeader must keep 16-byte alignment of the arrays");

// Classification grid (built once on the host, see api.cu): three u32 words per cell of a uniform
// grid over the polygon's bounding box (stored as three planes: word 0 of every cell, then word 1, then word 2, so
// the one word most waypoints need is dense in L1), so that most waypoints are classified by ONE load instead
// of the loops over circles / polygon edges / habitats / shark cells, and the rest by a short
// per-cell candidate list.  A code is only definitive when every point of the cell (plus a rounding
// margin) gets the same answer from the exact test; candidate lists hold every object that can
// decide a point of the cell.  Results are therefore identical to the exact tests by construction.
//  word 0  bits 0-1  polygon: 0 ambiguous, 1 strictly inside, 2 outside
//          bit  2    1: clear of every (inflated) obstacle circle
//          bits 3-10 habitat: 0..63 first-match habitat, 254 none, 255 ambiguous
//          bit  11   circles: more than 3 candidates -> full loop
//          bit  12   habitats: more than 3 candidates -> full loop
//          bit  13   polygon: convex fast path not applicable / more than 2 candidate edges -> full test
//          bits 16-31 shark cell: first-match cell id, 0xFFFF none, 0xFFFE ambiguous
//  word 1  three 10-bit circle indices (0x3FF = none): the only circles a point of the cell can hit
//  word 2  bits 0-17 three 6-bit habitat indices (0x3F = none), in list order; bits 18-27 two 5-bit
//          polygon edge indices (0x1F = none): the only edges whose half-plane is not already decided
#define AUV_GRID_HAB_NONE 254u
#define AUV_GRID_HAB_AMBIG 255u
#define AUV_GRID_CELL_NONE 0xFFFFu
#define AUV_GRID_CELL_AMBIG 0xFFFEu
#define AUV_GRID_CIRC_MANY (1u << 11)
#define AUV_GRID_HAB_MANY (1u << 12)
#define AUV_GRID_POLY_FULL (1u << 13)
#define AUV_GRID_ALL_AMBIG ((AUV_GRID_CELL_AMBIG << 16) | (AUV_GRID_HAB_AMBIG << 3) | AUV_GRID_CIRC_MANY | AUV_GRID_HAB_MANY | AUV_GRID_POLY_FULL)

struct Cls { unsigned code; int idx; };   // word 0 and the cell index (-1 outside the grid)

// Fine x table for the shark-cell lookup (cost.py:181-184, see the cell index above).  Bucket b covers
// [xc0 + b w, xc0 + (b+1) w]; when no breakpoint lies within a margin of it, every x of the bucket is in the
// same open piece and the entry holds that piece's FIRST candidate:
//   v >= 0 : cell id (bits 0-29) of the first candidate, c1 its lower y bound; bit 30: more candidates follow
//   v == -1: the piece has no candidate (no cell contains this x)
//   v == -2: a breakpoint is near: walk the breakpoints (find_cell)
template <typename R> struct alignas(8) XCell { R c1; int v; };

template <typename R> struct EnvView {
    int K, E, H, T, C, NB, NP, convex;
    int gnx, gny, ncell, bins_uniform, nxb, nxc;
    R gx0, gy0, ginv, gxo, gyo, bin_s0, bin_w, bin_winv, xb0, xbinv, xcinv, xco;
    const unsigned *grid;
    const XCell<R> *xcell;
    const unsigned short *xb;
    R minx, miny, maxx, maxy;
    const R *cx, *cy, *cr, *creff, *creff2;
    const R *px, *py;
    const R *hx, *hy, *hr, *hr2;
    const R *b0, *b1;
    const R *brk;
    const int *piece;
    const R *c1;
    const int *cell;
    const R *probs;

    __device__ __forceinline__ void bind(const unsigned char *hot, const unsigned char *probs_base) {
        const EnvHeader *h = (const EnvHeader *)hot;
        K = h->K; E = h->E; H = h->H; T = h->T; C = h->C; NB = h->NB; NP = h->NP; convex = h->convex;
        minx = (R)h->bbox[0]; miny = (R)h->bbox[1]; maxx = (R)h->bbox[2]; maxy = (R)h->bbox[3];
        cx = (const R *)(hot + h->off_cx); cy = (const R *)(hot + h->off_cy);
        cr = (const R *)(hot + h->off_cr); creff = (const R *)(hot + h->off_creff);
        creff2 = (const R *)(hot + h->off_creff2);
        px = (const R *)(hot + h->off_px); py = (const R *)(hot + h->off_py);
        hx = (const R *)(hot + h->off_hx); hy = (const R *)(hot + h->off_hy);
        hr = (const R *)(hot + h->off_hr); hr2 = (const R *)(hot + h->off_hr2);
        b0 = (const R *)(hot + h->off_b0); b1 = (const R *)(hot + h->off_b1);
        brk = (const R *)(hot + h->off_brk); piece = (const int *)(hot + h->off_piece);
        c1 = (const R *)(hot + h->off_c1); cell = (const int *)(hot + h->off_cell);
        probs = (const R *)(probs_base + h->off_probs);
    }
    // the grid stays in global memory: bind it from the blob in HBM
    __device__ __forceinline__ void bind_grid(const unsigned char *blob_global, const unsigned char *hot) {
        const EnvHeader *h = (const EnvHeader *)hot;
        gnx = h->gnx; gny = h->gny; bins_uniform = h->bins_uniform; nxb = h->nxb;
        xb0 = (R)h->xb0; xbinv = (R)(1.0 / h->xbw); xb = (const unsigned short *)(hot + h->off_xb);
        gx0 = (R)h->gx0; gy0 = (R)h->gy0; ginv = (R)(1.0 / h->gs);
        gxo = (R)(-h->gx0 / h->gs); gyo = (R)(-h->gy0 / h->gs);
        bin_s0 = (R)h->bin_s0; bin_w = (R)h->bin_w; bin_winv = (R)(1.0 / h->bin_w);
        grid = (const unsigned *)(blob_global + h->off_grid);
        ncell = gnx * gny;
        nxc = h->nxc; xcinv = (R)(1.0 / h->xcw); xco = (R)(-h->xc0 / h->xcw);
        xcell = (const XCell<R> *)(blob_global + h->off_xcell);
    }
    // the fine x table staged in shared memory too (the kernel copied [0, ext_bytes))
    __device__ __forceinline__ void bind_xcell_staged(const unsigned char *hot) {
        const EnvHeader *h = (const EnvHeader *)hot;
        xcell = (const XCell<R> *)(hot + h->off_xcell);
    }
    // classification of the cell containing (x, y); off the grid: outside the polygon, everything else ambiguous
    __device__ __forceinline__ Cls classify(R x, R y) const {
        Cls c;
        int ix, iy;
        if (sizeof(R) == 4) {
            // floor via round-toward-minus-infinity onto 2^23: two full-rate instructions per coordinate instead of
            // F2I on the conversion pipe; anything out of range (negative, huge, NaN) fails the unsigned compares
            const float fx = fmaf((float)x, (float)ginv, (float)gxo), fy = fmaf((float)y, (float)ginv, (float)gyo);
            ix = __float_as_int(__fadd_rd(fx, 8388608.f)) - 0x4B000000;
            iy = __float_as_int(__fadd_rd(fy, 8388608.f)) - 0x4B000000;
        } else {
            const R fx = (x - gx0) * ginv, fy = (y - gy0) * ginv;
            const bool in = fx >= (R)0 && fy >= (R)0 && fx < (R)gnx && fy < (R)gny;
            ix = in ? (int)fx : -1; iy = in ? (int)fy : -1;
        }
        if (!((unsigned)ix < (unsigned)gnx && (unsigned)iy < (unsigned)gny)) {
            // the grid covers the polygon's bounding box plus one cell all round: a point off the grid is outside the
            // polygon (code 2) whatever else is undecided about it
            c.code = AUV_GRID_ALL_AMBIG | (gnx > 0 ? 2u : 0u); c.idx = -1; return c;
        }
        c.idx = iy * gnx + ix;
        c.code = __ldg(grid + c.idx);
        return c;
    }
    __device__ __forceinline__ unsigned word1(const Cls &c) const { return __ldg(grid + ncell + c.idx); }
    __device__ __forceinline__ unsigned word2(const Cls &c) const { return __ldg(grid + 2 * ncell + c.idx); }
};

// ---- TMA bulk staging -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// Stage `bytes` (multiple of 16) of the blob into shared memory.  One elected thread arms the
// mbarrier with the byte count and issues the bulk copies (<= 64 KiB each); everybody waits.
// Must be called by all threads of the CTA.
__device__ __forceinline__ void stage_env_tma(unsigned char *smem_dst, const unsigned char *blob,
                                              int bytes, uint64_t *bar) {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)bytes);
        for (int o = 0; o < bytes; o += 65536) {
            int n = bytes - o < 65536 ? bytes - o : 65536;
            tma_bulk_g2s(smem_dst + o, blob + o, (uint32_t)n, bar);
        }
    }
    mbar_wait(bar, 0);
}

// bytes of dynamic shared memory a kernel needs for the env given a budget
struct EnvStagePlan { int bytes; int probs_staged; };

}  // namespace auv
