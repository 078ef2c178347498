// occupancy.cu -- SharkOccupancyGrid.convert (/root/reference/path_planning/sharkOccupancyGrid.py:47-71):
// shark tracks -> per-time-bin AUV-detection probability grids, the [T][C] input of the path cost.
// the data format on the producer side of the hot path.
//
//   k_occ_hist  one thread per track point: first time bin containing t (:253-258), first cell (in
//               cell_list order) that contains the point, boundary included (`point.within(cell) or
//               cell.touches(point)`, :232; exact orientation predicate) -> integer histogram
//   k_occ_norm  occupancy of (bin, shark, square): 0.01 for squares that hold a cell, +1 per point
//               (added one by one, as the reference does), / (points + 0.01 * cells)   (:217-241)
//   k_occ_auv   disc sum of radius ceil(range / cell) around every cell, in the reference's window
//               order (:189-201)
//   k_occ_avg   mean over sharks, in dict order                                           (:160-171)
// fp64 with separately rounded operations: bit-identical to the reference (tests/golden/occupancy.npz).
#include <math.h>
#include <vector>
#include "launch.h"

namespace auv {

// +1 strictly inside, 0 on the boundary, -1 outside; ragged polygon, exact
__device__ __forceinline__ int poly_locate(const double *xy, int n, double px, double py) {
    bool inside = false, boundary = false;
    for (int i = 0; i < n; i++) {
        const int j = i + 1 == n ? 0 : i + 1;
        const double ax = xy[2 * i], ay = xy[2 * i + 1], bx = xy[2 * j], by = xy[2 * j + 1];
        if (px == ax && py == ay) boundary = true;
        if (ay == py && by == py) {
            if (fmin(ax, bx) <= px && px <= fmax(ax, bx)) boundary = true;
            continue;
        }
        if ((ay > py) != (by > py)) {
            const int s = orient2d(ax, ay, bx, by, px, py);
            if (s == 0) boundary = true;
            if ((s > 0) == (by > ay)) inside = !inside;
        }
    }
    return boundary ? 0 : (inside ? 1 : -1);
}

__global__ void __launch_bounds__(256) k_occ_hist(const double *cell_xy, const long long *cell_off, const double *cell_bb,
                                                  const int *cell_sq, int C, const double *trk, const int *trk_shark,
                                                  long long n, int T, int S, int RC, double bin_interval, int *counts,
                                                  int *npts) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double x = trk[3 * k], y = trk[3 * k + 1], t = trk[3 * k + 2];
    int tb = -1;
    for (int b = 0; b < T; b++)
        if (t >= __dmul_rn((double)b, bin_interval) && t <= __dmul_rn((double)(b + 1), bin_interval)) { tb = b; break; }
    if (tb < 0) return;
    const int s = trk_shark[k];
    atomicAdd(&npts[tb * S + s], 1);
    for (int c = 0; c < C; c++) {
        const double *bb = cell_bb + 4 * c;
        if (x < bb[0] || x > bb[2] || y < bb[1] || y > bb[3]) continue;       // outside the cell's bounding box
        if (poly_locate(cell_xy + 2 * cell_off[c], (int)(cell_off[c + 1] - cell_off[c]), x, y) >= 0) {
            atomicAdd(&counts[((size_t)tb * S + s) * RC + cell_sq[c]], 1);
            break;
        }
    }
}

__global__ void __launch_bounds__(256) k_occ_norm(const int *counts, const int *npts, const unsigned char *is_cell, int T,
                                                  int S, int RC, int C, double *occ) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)T * S * RC) return;
    const int sq = (int)(i % RC);
    const long long bs = i / RC;
    double v = is_cell[sq] ? 0.01 : 0.0;
    const int cnt = counts[i];
    for (int k = 0; k < cnt; k++) v = __dadd_rn(v, 1.0);                        // `grid[row][col] += 1` per point
    const double nor = __dadd_rn((double)npts[bs], __dmul_rn((double)C, 0.01));
    occ[i] = __ddiv_rn(v, nor);
}

__global__ void __launch_bounds__(256) k_occ_auv(const double *occ, const int *sq_off, int T, int S,
                                                 int rows, int cols, int count, double *auv) {
    const int RC = rows * cols;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)T * S * RC) return;
    const int sq = (int)(i % RC), row = sq / cols, col = sq % cols;
    const double *o = occ + (i - sq);
    double acc = 0.0;
    for (int cc = sq_off[sq]; cc < sq_off[sq + 1]; cc++) {        // one pass per cell mapped to this square (:190)
        for (int a = 0; a < 4 * count; a++) {
            const int rt = row - 2 * count + a;
            for (int b = 0; b < 4 * count; b++) {
                const int ct = col - 2 * count + b;
                if (rt >= 0 && rt < rows && ct >= 0 && ct < cols) {
                    const int dr = rt - row, dc = ct - col;
                    if (dr * dr + dc * dc <= count * count) acc = __dadd_rn(acc, o[rt * cols + ct]);
                }
            }
        }
    }
    auv[i] = acc;
}

__global__ void __launch_bounds__(256) k_occ_avg(const double *auv, int T, int S, int RC, double *grid) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)T * RC) return;
    const int sq = (int)(i % RC);
    const int b = (int)(i / RC);
    double acc = 0.0;
    for (int s = 0; s < S; s++) acc = __dadd_rn(acc, auv[((size_t)b * S + s) * RC + sq]);
    grid[i] = __ddiv_rn(acc, (double)S);
}

}  // namespace auv

using namespace auv;

extern "C" int auvrrt_occupancy_dims(const double bounds[4], double cell_size, double bin_interval, const double *tracks,
                                     const int64_t *track_off, int S, int *T, int *rows, int *cols) {
    if (!bounds || !track_off || !T || !rows || !cols || !(cell_size > 0) || !(bin_interval > 0))
        return set_err(AUVRRT_ERR_ARG, "occupancy_dims: bad argument");
    double longest = 0;                                                                  // createBinList :307-314
    for (int s = 0; s < S; s++)
        if (track_off[s + 1] > track_off[s] && tracks[3 * (track_off[s + 1] - 1) + 2] > longest)
            longest = tracks[3 * (track_off[s + 1] - 1) + 2];
    *T = (int)floor(longest / bin_interval);
    *cols = (int)(ceil(bounds[2] - bounds[0]) / cell_size) + 1;
    *rows = (int)(ceil(bounds[3] - bounds[1]) / cell_size) + 1;
    return AUVRRT_OK;
}

extern "C" int auvrrt_occupancy_grid(const double *cell_xy, const int64_t *cell_off, int C, const double bounds[4],
                                     double cell_size, double bin_interval, double detect_range, const double *tracks,
                                     const int64_t *track_off, int S, int device, double *out_grid, int64_t out_cap) {
    int T, rows, cols;
    int rc = auvrrt_occupancy_dims(bounds, cell_size, bin_interval, tracks, track_off, S, &T, &rows, &cols);
    if (rc) return rc;
    if (C < 0 || S <= 0 || !out_grid) return set_err(AUVRRT_ERR_ARG, "occupancy_grid: bad argument");
    const int RC = rows * cols;
    if ((int64_t)T * RC > out_cap) return set_err(AUVRRT_ERR_ARG, "occupancy_grid: output holds %lld values, %lld needed", (long long)out_cap, (long long)T * RC);
    if (T == 0) return AUVRRT_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return set_err(AUVRRT_ERR_CUDA, "no CUDA device: libauvrrt has no CPU fallback"); }
    AUV_CUDA(cudaSetDevice(device));
    // host-side flattening: bounding boxes, square index of every cell (cellToIndex :294-299), CSR square -> cells
    std::vector<double> bb(4 * (size_t)std::max(C, 1));
    std::vector<int> csq(std::max(C, 1)), sq_off(RC + 1, 0), sq_cells(std::max(C, 1));
    std::vector<unsigned char> is_cell(RC, 0);
    for (int c = 0; c < C; c++) {
        double x0 = INFINITY, y0 = INFINITY, x1 = -INFINITY, y1 = -INFINITY;
        for (int64_t k = cell_off[c]; k < cell_off[c + 1]; k++) {
            x0 = fmin(x0, cell_xy[2 * k]); x1 = fmax(x1, cell_xy[2 * k]); y0 = fmin(y0, cell_xy[2 * k + 1]); y1 = fmax(y1, cell_xy[2 * k + 1]);
        }
        bb[4 * c] = x0; bb[4 * c + 1] = y0; bb[4 * c + 2] = x1; bb[4 * c + 3] = y1;
        const int col = (int)((x0 - bounds[0]) / cell_size), row = (int)((y0 - bounds[1]) / cell_size);
        if (row < 0 || row >= rows || col < 0 || col >= cols) return set_err(AUVRRT_ERR_ARG, "occupancy_grid: cell %d lies outside the boundary's grid (IndexError in the reference)", c);
        csq[c] = row * cols + col; is_cell[csq[c]] = 1; sq_off[csq[c] + 1]++;
    }
    for (int i = 0; i < RC; i++) sq_off[i + 1] += sq_off[i];
    { std::vector<int> fill(sq_off.begin(), sq_off.end() - 1); for (int c = 0; c < C; c++) sq_cells[fill[csq[c]]++] = c; }
    const int64_t n = track_off[S];
    std::vector<int> shark_of((size_t)std::max<int64_t>(n, 1));
    for (int s = 0; s < S; s++) for (int64_t k = track_off[s]; k < track_off[s + 1]; k++) shark_of[k] = s;
    const int count = (int)ceil(detect_range / cell_size);

    struct Buf { void *p = nullptr; ~Buf() { if (p) cudaFree(p); } } d_xy, d_off, d_bb, d_csq, d_trk, d_sh, d_cnt, d_np, d_isc, d_occ, d_auv, d_grid, d_sqo, d_sqc;
    auto up = [&](Buf &b, const void *src, size_t bytes) -> int {
        AUV_CUDA(cudaMalloc(&b.p, bytes ? bytes : 16));
        if (bytes && src) AUV_CUDA(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
        return AUVRRT_OK;
    };
    const size_t nv = (size_t)cell_off[C];
    if ((rc = up(d_xy, cell_xy, 16 * nv)) || (rc = up(d_off, cell_off, 8 * (size_t)(C + 1))) || (rc = up(d_bb, bb.data(), 32 * (size_t)C)) ||
        (rc = up(d_csq, csq.data(), 4 * (size_t)C)) || (rc = up(d_trk, tracks, 24 * (size_t)n)) || (rc = up(d_sh, shark_of.data(), 4 * (size_t)n)) ||
        (rc = up(d_isc, is_cell.data(), (size_t)RC)) || (rc = up(d_sqo, sq_off.data(), 4 * (size_t)(RC + 1))) || (rc = up(d_sqc, sq_cells.data(), 4 * (size_t)C)) ||
        (rc = up(d_cnt, nullptr, 4 * (size_t)T * S * RC)) || (rc = up(d_np, nullptr, 4 * (size_t)T * S)) ||
        (rc = up(d_occ, nullptr, 8 * (size_t)T * S * RC)) || (rc = up(d_auv, nullptr, 8 * (size_t)T * S * RC)) || (rc = up(d_grid, nullptr, 8 * (size_t)T * RC)))
        return rc;
    AUV_CUDA(cudaMemset(d_cnt.p, 0, 4 * (size_t)T * S * RC));
    AUV_CUDA(cudaMemset(d_np.p, 0, 4 * (size_t)T * S));
    const long long tot = (long long)T * S * RC;
    if (n > 0) {
        k_occ_hist<<<(unsigned)((n + 255) / 256), 256>>>((const double *)d_xy.p, (const long long *)d_off.p, (const double *)d_bb.p,
                                                         (const int *)d_csq.p, C, (const double *)d_trk.p, (const int *)d_sh.p, n, T, S, RC,
                                                         bin_interval, (int *)d_cnt.p, (int *)d_np.p);
        g_launches++;
    }
    k_occ_norm<<<(unsigned)((tot + 255) / 256), 256>>>((const int *)d_cnt.p, (const int *)d_np.p, (const unsigned char *)d_isc.p, T, S, RC, C, (double *)d_occ.p);
    k_occ_auv<<<(unsigned)((tot + 255) / 256), 256>>>((const double *)d_occ.p, (const int *)d_sqo.p, T, S, rows, cols, count, (double *)d_auv.p);
    k_occ_avg<<<(unsigned)(((long long)T * RC + 255) / 256), 256>>>((const double *)d_auv.p, T, S, RC, (double *)d_grid.p);
    g_launches += 3;
    AUV_CUDA(cudaGetLastError());
    AUV_CUDA(cudaMemcpy(out_grid, d_grid.p, 8 * (size_t)T * RC, cudaMemcpyDeviceToHost));
    return AUVRRT_OK;
}
