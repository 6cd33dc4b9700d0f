// cck_tree.cuh — SURVEY §8(f) row 2: RealVectorBeliefSpace::distance (src/Spaces/RealVectorBeliefSpace.cpp:16-62)
// and the two nearest-neighbour queries the low-level planner issues every iteration,
// selectNode (ConstraintRespectingBSST.cpp:119-148) and findClosestWitness (:150-178),
// as brute-force batched reductions over a device-resident tree.
//
//   distance(s1, s2) = | ||mu1 - mu2||^2 + trace(C1 + C2 - 2 (C2^1/2 C1 C2^1/2)^1/2) |,   0 if the means coincide
//
// with C = Sigma + Lambda and ^1/2 the symmetric square root (SelfAdjointEigenSolver::operatorSqrt,
// which reads the lower triangle).  The eigen-decomposition is a cyclic Jacobi iteration on the
// symmetric part; the CPU oracle runs the same operation sequence, so distances and the selected
// indices are bit-identical to it (Eigen's tridiagonal QL agrees to rounding).
//
// Layout: items (tree nodes or witnesses) are SoA planes of `cap` doubles:
//   mean[DIM] | cov[DIM*DIM] | sqrtcov[DIM*DIM] | cost ; plus an `active` byte plane.
// sqrtcov is filled once per item at append time.  One CTA scans all items for one query; lanes
// reduce (key, index) pairs lexicographically so ties resolve to the lowest index, exactly like the
// sequential scan.  A NaN distance (a rank-deficient covariance whose product gets a slightly negative eigenvalue under
// operatorSqrt, or non-finite inputs) has no defined answer: a sequential scan keeps its first item if THAT distance is
// NaN and the reference's GNAT is undefined; distances themselves stay bit-identical to the oracle, NaN included
// (tests/hostsim/fuzz_rollout.py --tree).
#pragma once
#include "cck_device.cuh"

namespace cck {

// symmetric square root of the lower-triangle-symmetrised A (DIM x DIM): out = V diag(sqrt(l)) V^T (NaN for an indefinite input, as Eigen's SelfAdjointEigenSolver::operatorSqrt).
// DIAG_ONLY: only out[i*DIM+i] is produced (that is all the trace needs).
template <int DIM, bool DIAG_ONLY>
__host__ __device__ inline void sym_sqrt(const double* Ain, double* out) {
    double A[DIM * DIM], V[DIM * DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
            A[i * DIM + j] = Ain[(i > j ? i : j) * DIM + (i > j ? j : i)];
            V[i * DIM + j] = i == j ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int j = 0; j < DIM; ++j)
                if (j < i) off += A[i * DIM + j] * A[i * DIM + j];
        // converged: the off-diagonal is at the rounding floor (<= 1e-15 of the diagonal in norm).  Rounding leaves a residue
        // of ~1e-17 there that never reaches the absolute 1e-300 test, so that test alone ran all 60 sweeps every time
        // (ncu: 9000 instructions per warp for one 2x2 distance) - ten times the work of the converged iteration.
        double dg = 0;
#pragma unroll
        for (int i = 0; i < DIM; ++i) dg += A[i * DIM + i] * A[i * DIM + i];
        if (off < 1e-300 || off <= 1e-30 * dg) break;
#pragma unroll
        for (int p = 0; p < DIM; ++p)
#pragma unroll
            for (int q = 0; q < DIM; ++q) {
                if (q <= p) continue;
                if (fabs(A[p * DIM + q]) < 1e-300) continue;
                const double th = (A[q * DIM + q] - A[p * DIM + p]) / (2 * A[p * DIM + q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
                const double c = 1 / sqrt(t * t + 1), s = t * c;
#pragma unroll
                for (int k = 0; k < DIM; ++k) { const double akp = A[k * DIM + p], akq = A[k * DIM + q]; A[k * DIM + p] = c * akp - s * akq; A[k * DIM + q] = s * akp + c * akq; }
#pragma unroll
                for (int k = 0; k < DIM; ++k) { const double apk = A[p * DIM + k], aqk = A[q * DIM + k]; A[p * DIM + k] = c * apk - s * aqk; A[q * DIM + k] = s * apk + c * aqk; }
#pragma unroll
                for (int k = 0; k < DIM; ++k) { const double vkp = V[k * DIM + p], vkq = V[k * DIM + q]; V[k * DIM + p] = c * vkp - s * vkq; V[k * DIM + q] = s * vkp + c * vkq; }
            }
    }
    double sq[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) sq[k] = sqrt(A[k * DIM + k]);   // Eigen's operatorSqrt is eigenvalues().cwiseSqrt(): a negative eigenvalue (indefinite lower triangle) gives NaN, not 0
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
            if (DIAG_ONLY && i != j) continue;
            double acc = 0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) acc += V[i * DIM + k] * sq[k] * V[j * DIM + k];
            out[i * DIM + j] = acc;
        }
}

// trace(C1 + C2 - 2 (S2 C1 S2)^1/2) with S2 = C2^1/2 given
template <int DIM>
__host__ __device__ inline double cov_term(const double* C1, const double* C2, const double* S2) {
    double t[DIM * DIM], m[DIM * DIM], s3[DIM * DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
            double a = 0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) a += S2[i * DIM + k] * C1[k * DIM + j];
            t[i * DIM + j] = a;
        }
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) {
            double a = 0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) a += t[i * DIM + k] * S2[k * DIM + j];
            m[i * DIM + j] = a;
        }
    sym_sqrt<DIM, true>(m, s3);
    double tr = 0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) tr += C1[i * DIM + i] + C2[i * DIM + i] - 2 * s3[i * DIM + i];
    return tr;
}

struct TreeDev {
    double* planes;      // [DIM + 2*DIM*DIM + 1][cap]
    uint8_t* active;     // [cap]
    int64_t cap;
};
template <int DIM> __host__ __device__ constexpr int tree_planes() { return DIM + 2 * DIM * DIM + 1; }

// fill cov / sqrtcov / cost / active of items [first, first+n) from SoA belief records (device)
template <int DIM>
__global__ void __launch_bounds__(128) k_tree_append(TreeDev tr, int64_t first, int64_t n, const double* __restrict__ bel, int64_t ld,
                                                     const double* __restrict__ cost) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = first + j;
        double C[DIM * DIM], S[DIM * DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) tr.planes[d * tr.cap + i] = bel[d * ld + j];
#pragma unroll
        for (int f = 0; f < DIM * DIM; ++f) C[f] = bel[(DIM + f) * ld + j] + bel[(DIM + DIM * DIM + f) * ld + j];
        sym_sqrt<DIM, false>(C, S);
#pragma unroll
        for (int f = 0; f < DIM * DIM; ++f) { tr.planes[(DIM + f) * tr.cap + i] = C[f]; tr.planes[(DIM + DIM * DIM + f) * tr.cap + i] = S[f]; }
        tr.planes[(DIM + 2 * DIM * DIM) * tr.cap + i] = cost[j];
        tr.active[i] = 1;
    }
}

__global__ void __launch_bounds__(128) k_tree_update(TreeDev tr, int planes_cost, int64_t n, const int32_t* __restrict__ idx,
                                                     const double* __restrict__ cost, const uint8_t* __restrict__ active) {
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx[j];
        if (cost) tr.planes[(int64_t)planes_cost * tr.cap + i] = cost[j];
        if (active) tr.active[i] = active[j];
    }
}

struct KeyIdx { double key; long long idx; };
__device__ __forceinline__ void keyidx_min(KeyIdx& a, double k, long long i) {
    if (i >= 0 && (a.idx < 0 || k < a.key || (k == a.key && i < a.idx))) { a.key = k; a.idx = i; }
}
__device__ __forceinline__ void keyidx_warp(KeyIdx& a) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double k = __shfl_xor_sync(0xffffffffu, a.key, o);
        const long long i = __shfl_xor_sync(0xffffffffu, a.idx, o);
        keyidx_min(a, k, i);
    }
}

// MODE 0 (selectNode): query = sample, item = tree node; distance(sample, node) uses the NODE's sqrt
//   (cov_term(C_sample, C_node)); among active nodes: best cost within `radius`, else nearest.
// MODE 1 (findClosestWitness): query = new node, item = witness; cov_term(C_witness, C_query) uses the
//   QUERY's sqrt; plain nearest, no active filter.
// One CTA per query.  out_idx[q] = chosen item or -1, out_dist[q] = its distance.
template <int DIM, int MODE>
__global__ void __launch_bounds__(128) k_tree_query(TreeDev tr, int64_t n_items, const double* __restrict__ qbel, int64_t ldq, double radius,
                                                    int32_t* __restrict__ out_idx, double* __restrict__ out_dist) {
    const int64_t q = blockIdx.x;
    double qm[DIM], qC[DIM * DIM], qS[DIM * DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) qm[d] = qbel[d * ldq + q];
#pragma unroll
    for (int f = 0; f < DIM * DIM; ++f) qC[f] = qbel[(DIM + f) * ldq + q] + qbel[(DIM + DIM * DIM + f) * ldq + q];
    if (MODE == 1) sym_sqrt<DIM, false>(qC, qS);
    KeyIdx best{0.0, -1}, near{0.0, -1}, bestd{0.0, -1};
    for (int64_t i = threadIdx.x; i < n_items; i += blockDim.x) {
        if (MODE == 0 && !tr.active[i]) continue;
        double d2 = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { const double e = qm[d] - tr.planes[d * tr.cap + i]; d2 += e * e; }
        double dist = 0;
        if (d2 != 0) {
            double C[DIM * DIM], S[DIM * DIM];
#pragma unroll
            for (int f = 0; f < DIM * DIM; ++f) C[f] = tr.planes[(DIM + f) * tr.cap + i];
            if (MODE == 0) {
#pragma unroll
                for (int f = 0; f < DIM * DIM; ++f) S[f] = tr.planes[(DIM + DIM * DIM + f) * tr.cap + i];
                dist = fabs(d2 + cov_term<DIM>(qC, C, S));
            } else {
                dist = fabs(d2 + cov_term<DIM>(C, qC, qS));
            }
        }
        if (MODE == 0 && dist <= radius) {
            const double cst = tr.planes[(DIM + 2 * DIM * DIM) * tr.cap + i];
            if (best.idx < 0 || cst < best.key) { best.key = cst; best.idx = i; bestd.key = dist; bestd.idx = i; }
        }
        if (near.idx < 0 || dist < near.key) { near.key = dist; near.idx = i; }
    }
    // block reduction (lexicographic (key, index): ties -> lowest index, as the sequential scan)
    __shared__ double s_key[2][4];
    __shared__ long long s_idx[2][4];
    __shared__ double s_bd[4];
    keyidx_warp(near);
    // for `best` the distance of the winner travels with it
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double k = __shfl_xor_sync(0xffffffffu, best.key, o);
            const long long i = __shfl_xor_sync(0xffffffffu, best.idx, o);
            const double dd = __shfl_xor_sync(0xffffffffu, bestd.key, o);
            if (i >= 0 && (best.idx < 0 || k < best.key || (k == best.key && i < best.idx))) { best.key = k; best.idx = i; bestd.key = dd; }
        }
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_key[0][w] = near.key; s_idx[0][w] = near.idx; s_key[1][w] = best.key; s_idx[1][w] = best.idx; s_bd[w] = bestd.key; }
    __syncthreads();
    if (threadIdx.x == 0) {
        KeyIdx n2{0.0, -1}, b2{0.0, -1};
        double bd = 0;
        for (int k = 0; k < 4; ++k) {
            keyidx_min(n2, s_key[0][k], s_idx[0][k]);
            const long long bi = s_idx[1][k];
            if (bi >= 0 && (b2.idx < 0 || s_key[1][k] < b2.key || (s_key[1][k] == b2.key && bi < b2.idx))) { b2.key = s_key[1][k]; b2.idx = bi; bd = s_bd[k]; }
        }
        if (MODE == 0 && b2.idx >= 0) { out_idx[q] = (int32_t)b2.idx; if (out_dist) out_dist[q] = bd; }
        else { out_idx[q] = (int32_t)n2.idx; if (out_dist) out_dist[q] = n2.idx >= 0 ? n2.key : HUGE_VAL; }
    }
}

// lower triangle of the pairwise matrix of one batch: out[q*n + i] = distance(b_i, b_q) for i < q
// (findClosestWitness between candidates of the same launch: witness = earlier candidate i, node = q)
template <int DIM>
__global__ void __launch_bounds__(128) k_distance_lower(int64_t n, const double* __restrict__ bel, int64_t ld, double* __restrict__ out) {
    const int64_t q = blockIdx.x;
    double qm[DIM], qC[DIM * DIM], qS[DIM * DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) qm[d] = bel[d * ld + q];
#pragma unroll
    for (int f = 0; f < DIM * DIM; ++f) qC[f] = bel[(DIM + f) * ld + q] + bel[(DIM + DIM * DIM + f) * ld + q];
    sym_sqrt<DIM, false>(qC, qS);
    for (int64_t i = threadIdx.x; i < q; i += blockDim.x) {
        double d2 = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { const double e = bel[d * ld + i] - qm[d]; d2 += e * e; }
        double dist = 0;
        if (d2 != 0) {
            double C[DIM * DIM];
#pragma unroll
            for (int f = 0; f < DIM * DIM; ++f) C[f] = bel[(DIM + f) * ld + i] + bel[(DIM + DIM * DIM + f) * ld + i];
            dist = fabs(d2 + cov_term<DIM>(C, qC, qS));
        }
        out[q * n + i] = dist;
    }
}

// selected columns of the same matrix: out[q*nf + a] = distance(b_cols[a], b_q) for cols[a] < q.  Only candidates
// that are farther than the pruning radius from every existing witness can become witnesses themselves, so the
// planner asks for those columns only (usually a handful once the tree is established).
template <int DIM>
__global__ void __launch_bounds__(128) k_distance_cols(int64_t n, const double* __restrict__ bel, int64_t ld, int nf,
                                                       const int32_t* __restrict__ cols, double* __restrict__ out) {
    const int64_t q = blockIdx.x;
    double qm[DIM], qC[DIM * DIM], qS[DIM * DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) qm[d] = bel[d * ld + q];
#pragma unroll
    for (int f = 0; f < DIM * DIM; ++f) qC[f] = bel[(DIM + f) * ld + q] + bel[(DIM + DIM * DIM + f) * ld + q];
    sym_sqrt<DIM, false>(qC, qS);
    for (int a = threadIdx.x; a < nf; a += blockDim.x) {
        const int64_t i = cols[a];
        if (i >= q) continue;
        double d2 = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { const double e = bel[d * ld + i] - qm[d]; d2 += e * e; }
        double dist = 0;
        if (d2 != 0) {
            double C[DIM * DIM];
#pragma unroll
            for (int f = 0; f < DIM * DIM; ++f) C[f] = bel[(DIM + f) * ld + i] + bel[(DIM + DIM * DIM + f) * ld + i];
            dist = fabs(d2 + cov_term<DIM>(C, qC, qS));
        }
        out[q * nf + a] = dist;
    }
}

// The same matrix, but only the entries the sequential witness replay can use: for row q the pairs (a, distance) with
// cols[a] < q and distance < limit[q], in increasing a, at most `cap` of them (count[q] is the full number, so the
// caller sees an overflow).  One CTA of 128 threads per row (a distance is a long dependent chain of fp64 div/sqrt, so
// the columns of a row are spread over as many threads as possible); ordered compaction with ballot + popc per warp
// and a 4-entry prefix over the warps.  cols must be ascending.
template <int DIM>
__global__ void __launch_bounds__(128) k_distance_cols_near(int64_t n, const double* __restrict__ bel, int64_t ld, int nf,
                                                            const int32_t* __restrict__ cols, const double* __restrict__ limit, int cap,
                                                            int32_t* __restrict__ out_idx, double* __restrict__ out_dist, int32_t* __restrict__ count) {
    __shared__ int wcnt[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t q = blockIdx.x;
    double qm[DIM], qC[DIM * DIM], qS[DIM * DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) qm[d] = bel[d * ld + q];
#pragma unroll
    for (int f = 0; f < DIM * DIM; ++f) qC[f] = bel[(DIM + f) * ld + q] + bel[(DIM + DIM * DIM + f) * ld + q];
    sym_sqrt<DIM, false>(qC, qS);
    const double lim = limit[q];
    int base = 0;
    for (int a0 = 0; a0 < nf; a0 += 128) {
        if (cols[a0] >= q) break;                      // ascending: nothing further is below q (CTA-uniform)
        const int a = a0 + threadIdx.x;
        bool keep = false;
        double dist = 0;
        if (a < nf) {
            const int64_t i = cols[a];
            if (i < q) {
                double d2 = 0;
#pragma unroll
                for (int d = 0; d < DIM; ++d) { const double e = bel[d * ld + i] - qm[d]; d2 += e * e; }
                if (d2 != 0) {
                    double C[DIM * DIM];
#pragma unroll
                    for (int f = 0; f < DIM * DIM; ++f) C[f] = bel[(DIM + f) * ld + i] + bel[(DIM + DIM * DIM + f) * ld + i];
                    dist = fabs(d2 + cov_term<DIM>(C, qC, qS));
                }
                keep = dist < lim;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) wcnt[warp] = __popc(m);
        __syncthreads();
        int before = base, total = 0;
#pragma unroll
        for (int w = 0; w < 4; ++w) { if (w < warp) before += wcnt[w]; total += wcnt[w]; }
        if (keep) {
            const int pos = before + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) { out_idx[q * cap + pos] = a; out_dist[q * cap + pos] = dist; }
        }
        base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) count[q] = base;
}

// ---------------------------------------------------------------------------
// The witness rules of one launch, resolved on the device (ConstraintRespectingBSST.cpp:150-178 and :250-326).
// The reference handles the survivors one at a time; survivor s becomes a NEW witness iff no witness - existing or created
// by an earlier survivor of the same launch - lies within the pruning radius.  With the nearest EXISTING witness known
// (k_tree_query MODE 1), the same-launch part is a greedy scan over the "potential" survivors (no existing witness within
// the radius, and allowed to persist: `eligible`, i.e. not rejected by the constraints) in order:
//   fresh[t] = potential[t] && no fresh f < t with distance(f, t) <= radius.
// k_wit_rows forms the lower-triangular distances to potential columns and their "<= radius" bit rows, k_wit_greedy runs
// the scan from shared memory (one warp, one word of the fresh set per lane, n <= 1024), k_wit_rowmin gives every survivor
// its nearest fresh predecessor below its limit.  The host replays the remaining bookkeeping with O(1) work per survivor.
// ---------------------------------------------------------------------------
#define CCK_WIT_MAX 1024
__global__ void k_wit_prep(int n, const int32_t* __restrict__ widx, const double* __restrict__ wdist, double radius, double radius_up,
                           const uint8_t* __restrict__ eligible, uint8_t* __restrict__ iscol, uint8_t* __restrict__ cand, uint32_t* __restrict__ cand_words,
                           double* __restrict__ lim) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;     // grid covers CCK_WIT_MAX: every word of cand_words is written
    bool is_cand = false;
    if (s < n) {
        const bool far = widx[s] < 0 || wdist[s] > radius;
        iscol[s] = far ? 1 : 0;                                    // would be a new witness if no same-launch witness is near
        is_cand = far && (!eligible || eligible[s]);               // ... and would persist
        cand[s] = is_cand ? 1 : 0;
        const double w = widx[s] < 0 ? HUGE_VAL : wdist[s];
        lim[s] = w < radius_up ? w : radius_up;                    // only a same-launch witness nearer than this can matter
    }
    const unsigned m = __ballot_sync(0xffffffffu, is_cand);
    if ((threadIdx.x & 31) == 0) cand_words[s >> 5] = m;
}

// one CTA (128 threads) per survivor s: D[s][t] = distance(t, s) for candidate columns t < s (the survivor's own sqrt, like
// findClosestWitness), adj[s][t/32] bit t%32 = D[s][t] <= radius.  Words at and beyond s are written as zero.
template <int DIM>
__global__ void __launch_bounds__(128) k_wit_rows(int n, const double* __restrict__ bel, int64_t ld, const uint8_t* __restrict__ cand, double radius,
                                                  double* __restrict__ D, uint32_t* __restrict__ adj) {
    __shared__ double sq[DIM * DIM];
    const int s = blockIdx.x, lane = threadIdx.x & 31;
    double qm[DIM], qC[DIM * DIM], qS[DIM * DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) qm[d] = bel[d * ld + s];
#pragma unroll
    for (int f = 0; f < DIM * DIM; ++f) qC[f] = bel[(DIM + f) * ld + s] + bel[(DIM + DIM * DIM + f) * ld + s];
    if (threadIdx.x == 0) {          // the row's own square root once, not once per thread
        sym_sqrt<DIM, false>(qC, qS);
#pragma unroll
        for (int f = 0; f < DIM * DIM; ++f) sq[f] = qS[f];
    }
    __syncthreads();
#pragma unroll
    for (int f = 0; f < DIM * DIM; ++f) qS[f] = sq[f];
    const int nw = CCK_WIT_MAX / 32;
    for (int t0 = 0; t0 < nw * 32; t0 += 128) {
        const int t = t0 + threadIdx.x;
        bool near = false;
        if (t < s && cand[t]) {
            double d2 = 0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) { const double e = bel[d * ld + t] - qm[d]; d2 += e * e; }
            double dist = 0;
            if (d2 != 0) {
                double C[DIM * DIM];
#pragma unroll
                for (int f = 0; f < DIM * DIM; ++f) C[f] = bel[(DIM + f) * ld + t] + bel[(DIM + DIM * DIM + f) * ld + t];
                dist = fabs(d2 + cov_term<DIM>(C, qC, qS));
            }
            D[(size_t)s * n + t] = dist;
            near = dist <= radius;
        }
        const unsigned m = __ballot_sync(0xffffffffu, near);
        if (lane == 0) adj[(size_t)s * nw + (t >> 5)] = m;
    }
}

// one CTA: the bit rows go to shared memory (128 KB for 1024 survivors), warp 0 runs the sequential scan
__global__ void __launch_bounds__(1024) k_wit_greedy(int n, const uint32_t* __restrict__ cand_words, const uint32_t* __restrict__ adj,
                                                     uint32_t* __restrict__ fresh_words, uint8_t* __restrict__ fresh) {
    extern __shared__ uint32_t sadj[];
    const int nw = CCK_WIT_MAX / 32;
    for (int i = threadIdx.x; i < n * nw; i += blockDim.x) sadj[i] = adj[i];
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const uint32_t my_cand = cand_words[lane];   // word `lane` of the candidate set
    uint32_t mine = 0;                           // word `lane` of the fresh set
    for (int w = 0; w < (n + 31) / 32; ++w) {
        uint32_t cw = __shfl_sync(0xffffffffu, my_cand, w);
        if (!cw) continue;
        int b = __ffs(cw) - 1;
        uint32_t row = sadj[(w * 32 + b) * nw + lane];
        while (true) {
            cw &= cw - 1;
            const int bn = cw ? __ffs(cw) - 1 : b;
            const uint32_t row_next = sadj[(w * 32 + bn) * nw + lane];     // the next row is read before this one is voted on
            if (!__any_sync(0xffffffffu, (row & mine) != 0) && lane == w) mine |= 1u << b;
            if (!cw) break;
            b = bn; row = row_next;
        }
    }
    fresh_words[lane] = mine;
    for (int w = 0; w < nw; ++w) {
        const uint32_t word = __shfl_sync(0xffffffffu, mine, w);
        const int t = w * 32 + lane;
        if (t < n) fresh[t] = (word >> lane) & 1u;
    }
}

// one warp per survivor: nearest fresh predecessor with distance < lim[s]; ties resolve to the lowest index (the
// sequential scan keeps the first strict minimum)
__global__ void __launch_bounds__(128) k_wit_rowmin(int n, const double* __restrict__ D, const uint32_t* __restrict__ fresh_words, const double* __restrict__ lim,
                                                    int32_t* __restrict__ near_col, double* __restrict__ near_dist) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= n) return;
    const double l = lim[s];
    double bd = HUGE_VAL;
    int bt = -1;
    for (int t = lane; t < s; t += 32) {
        if (!((fresh_words[t >> 5] >> (t & 31)) & 1u)) continue;
        const double d = D[(size_t)s * n + t];
        if (d < l && d < bd) { bd = d; bt = t; }          // t ascends per lane: strict < keeps the lowest index
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double od = __shfl_down_sync(0xffffffffu, bd, o);
        const int ot = __shfl_down_sync(0xffffffffu, bt, o);
        if (ot >= 0 && (bt < 0 || od < bd || (od == bd && ot < bt))) { bd = od; bt = ot; }
    }
    if (lane == 0) { near_col[s] = bt; near_dist[s] = bt >= 0 ? bd : 0.0; }
}

// plain batched distance(a_i, b_i) (second argument's sqrt, as RealVectorBeliefSpace::distance(state1, state2))
template <int DIM>
__global__ void __launch_bounds__(128) k_belief_distance(int64_t n, const double* __restrict__ a, int64_t lda, const double* __restrict__ b,
                                                         int64_t ldb, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double d2 = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) { const double e = a[d * lda + i] - b[d * ldb + i]; d2 += e * e; }
        double dist = 0;
        if (d2 != 0) {
            double C1[DIM * DIM], C2[DIM * DIM], S2[DIM * DIM];
#pragma unroll
            for (int f = 0; f < DIM * DIM; ++f) {
                C1[f] = a[(DIM + f) * lda + i] + a[(DIM + DIM * DIM + f) * lda + i];
                C2[f] = b[(DIM + f) * ldb + i] + b[(DIM + DIM * DIM + f) * ldb + i];
            }
            sym_sqrt<DIM, false>(C2, S2);
            dist = fabs(d2 + cov_term<DIM>(C1, C2, S2));
        }
        out[i] = dist;
    }
}

}  // namespace cck
