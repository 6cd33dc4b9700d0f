// cck_b200.cu — implementation of the C ABI declared in include/cck_b200.h.
// Built with: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo ...
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "cck_kernels.cuh"
#include "cck_rollout_uni.cuh"
#include "cck_tree.cuh"

using namespace cck;

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
#define CCK_SLOTS 4   // pipeline slots of the host-pointer entry points (H2D / kernel / D2H of different chunks overlap)
struct cck_ctx {
    int device = 0;
    int num_sms = 0;
    int max_smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t pipe[CCK_SLOTS] = {};
    std::string err;
    int64_t launches = 0;

    bool has_model = false;
    cck_model_params model{};
    LinDev lin{};
    UniDev uni{};

    bool has_svc = false;
    SvcDev svc{};
    unsigned char* d_tab = nullptr;
    int tab_bytes = 0;
    uint32_t* d_df = nullptr;          // clearance field (first test of a belief-step, cck_rollout_grid.cuh), CCK_DF_N^2 words
    unsigned char* d_gtab = nullptr;   // fp32 box + cell-grid table (Blackmore broad phase, cck_rollout_grid.cuh)
    int gtab_bytes = 0;
    double* d_gexB = nullptr;          // exact fp64 rows in grid-table order
    int32_t* d_gexCls = nullptr;       // class of each row
    bool uni_sparse = false;           // unicycle matrices have the reference's exact-zero pattern
    GridConst gconst{0, 0, 0, 0, 0, 0, 0, -1.f};   // value-initialised tail: no clearance field   // header scalars of the grid table + fp32 bounds (kernel constants)
    int M = 0;

    bool has_pvc = false;
    cck_pvc_params pvc_host{};
    std::vector<double> pvc_A, pvc_B, pvc_radii;
    std::vector<int32_t> pvc_order;
    PvcDev pvc{};
    double* d_pvc_A = nullptr; double* d_pvc_B = nullptr; double* d_pvc_radii = nullptr; int32_t* d_pvc_order = nullptr;
    unsigned long long* d_key = nullptr;
    cck_conflict_run* d_run = nullptr;

    double* d_cov_tab = nullptr;       // covariance table of the planner's tree (cck_set_cov_table)
    int cov_tab_n = 0;
    unsigned char* d_cons = nullptr;   // constraint table (nullptr = no constraints)
    size_t cons_cap = 0;
    int cons_bytes = 0;                // size of the current constraint table (staged into shared memory when small)

    // pinned staging block of the small-batch path (planner-sized launches: one H2D, one launch, one D2H)
    unsigned char* h_stage = nullptr;
    size_t h_stage_cap = 0;
    // grow-only device scratch for the host-pointer entry points (one per pipeline slot)
    void* d_scratch[CCK_SLOTS] = {};
    size_t scratch_cap[CCK_SLOTS] = {};
};

static std::string g_create_err;
static int64_t kChunk = 1 << 20;   // beliefs per pipeline chunk of the host-pointer entry points (CCK_CHUNK_LOG2 overrides, tuning only)

#define CK(ctx, call)                                                                                  \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
            return CCK_ERR_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

static cudaError_t create_pipe_streams(cck_ctx* c) {
    for (int i = 0; i < CCK_SLOTS; ++i) {
        const cudaError_t e = cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static int fail(cck_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

extern "C" int cck_ctx_create(int device, cck_ctx** out) {
    if (!out) return CCK_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                       " (libcck_b200 has no CPU fallback)";
        return CCK_ERR_NODEVICE;
    }
    if (device < 0 || device >= ndev) { g_create_err = "device index out of range"; return CCK_ERR_INVALID; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return CCK_ERR_CUDA; }
    if (prop.major != 10) {
        g_create_err = "device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                       "; this library ships sm_100a code only (no fallback)";
        return CCK_ERR_NODEVICE;
    }
    cck_ctx* c = new cck_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    if (const char* e5 = getenv("CCK_CHUNK_LOG2")) { const int v = atoi(e5); if (v >= 10 && v <= 26) kChunk = (int64_t)1 << v; }
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = create_pipe_streams(c)) != cudaSuccess ||
        (e = cudaMalloc(&c->d_key, sizeof(unsigned long long))) != cudaSuccess || (e = cudaMalloc(&c->d_run, sizeof(cck_conflict_run))) != cudaSuccess) {
        g_create_err = cudaGetErrorString(e);
        delete c;
        return CCK_ERR_CUDA;
    }
    *out = c;
    return CCK_OK;
}

extern "C" int cck_ctx_destroy(cck_ctx* c) {
    if (!c) return CCK_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFreeHost(c->h_stage);
    for (int i = 0; i < CCK_SLOTS; ++i) { if (c->pipe[i]) { cudaStreamSynchronize(c->pipe[i]); cudaStreamDestroy(c->pipe[i]); } cudaFree(c->d_scratch[i]); }
    cudaFree(c->d_df); cudaFree(c->d_gtab); cudaFree(c->d_gexB); cudaFree(c->d_gexCls);
    cudaFree(c->d_tab); cudaFree(c->d_pvc_A); cudaFree(c->d_pvc_B); cudaFree(c->d_pvc_radii); cudaFree(c->d_pvc_order);
    cudaFree(c->d_key); cudaFree(c->d_run); cudaFree(c->d_cons); cudaFree(c->d_cov_tab);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return CCK_OK;
}

extern "C" const char* cck_last_error(const cck_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }
extern "C" int cck_num_sms(const cck_ctx* c) { return c ? c->num_sms : 0; }
extern "C" int64_t cck_launch_count(const cck_ctx* c) { return c ? c->launches : 0; }
extern "C" int cck_ctx_synchronize(cck_ctx* c) {
    if (!c) return CCK_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < CCK_SLOTS; ++i) CK(c, cudaStreamSynchronize(c->pipe[i]));
    return CCK_OK;
}

// ---------------------------------------------------------------------------
// host-side constant builders (restating the reference's constructor arithmetic)
// ---------------------------------------------------------------------------
// boost::math::erf_inv: Halley iterations on erfl/erfcl in long double (Boost promotes
// double to long double; the result is rounded to double).
extern "C" double cck_erf_inv(double z) {
    if (std::isnan(z) || z < -1.0 || z > 1.0) return std::nan("");
    if (z == 1.0) return HUGE_VAL;
    if (z == -1.0) return -HUGE_VAL;
    if (z == 0.0) return 0.0;
    const bool neg = z < 0;
    const long double p = neg ? -(long double)z : (long double)z;
    const long double q = 1.0L - p;
    const long double lg = logl(1.0L - p * p);
    const long double aw = 0.147L, pi = 3.14159265358979323846264338327950288L;
    const long double t1 = 2.0L / (pi * aw) + lg / 2.0L;
    long double x = sqrtl(sqrtl(t1 * t1 - lg / aw) - t1);
    const long double k = 1.12837916709551257389615890312154517L;   // 2/sqrt(pi)
    for (int it = 0; it < 60; ++it) {
        const long double f = (p > 0.5L) ? (q - erfcl(x)) : (erfl(x) - p);
        const long double d = k * expl(-x * x);
        const long double dx = f / (d + x * f);
        x -= dx;
        if (fabsl(dx) <= 1e-21L * fabsl(x)) break;
    }
    const double r = (double)x;
    return neg ? -r : r;
}
extern "C" double cck_chi2_quantile_2dof(double p) { return -2.0 * std::log1p(-p); }

extern "C" void cck_split_risk(double p_safe, int n_obstacles, int n_robots, double* p_safe_obs, double* p_safe_agents) {
    const double total_things = (double)(n_obstacles + n_robots);          // Instance.cpp:38
    const double p_coll = 1 - p_safe;
    const double p_coll_obs = (p_coll * n_obstacles) / total_things;
    const double p_coll_agnts = (p_coll * n_robots) / total_things;
    if (p_safe_agents) *p_safe_agents = 1 - p_coll_agnts;
    if (p_safe_obs) *p_safe_obs = 1 - p_coll_obs;
}

namespace {
struct P2 { double x, y; };
typedef std::vector<P2> Ring;
Ring rect_ring(double x, double y, double len, double width) {            // common.cpp:25-42
    P2 bl{x - (len / 2), y - (width / 2)}, br{x + (len / 2), y - (width / 2)};
    P2 tl{x - (len / 2), y + (width / 2)}, tr{x + (len / 2), y + (width / 2)};
    return Ring{bl, br, tr, tl, bl};
}
Ring square_ring(double r) { return Ring{{-r, -r}, {r, -r}, {r, r}, {-r, r}, {-r, -r}}; }
Ring minkowski(const Ring& a, const Ring& b) {                             // PCCBlackmoreSVC.cpp:110-145
    Ring s1 = a, s2 = b;
    s1.push_back(s1[1]); s2.push_back(s2[1]);
    Ring res;
    size_t i = 0, j = 0;
    while ((i < s1.size() - 2 || j < s2.size() - 2) && res.size() < 64) {
        res.push_back(P2{s1[i].x + s2[j].x, s1[i].y + s2[j].y});
        const double ax = s1[i + 1].x - s1[i].x, ay = s1[i + 1].y - s1[i].y;
        const double bx = s2[j + 1].x - s2[j].x, by = s2[j + 1].y - s2[j].y;
        const double cross = (ax * by) - (ay * bx);
        if (cross >= 0) ++i;
        if (cross <= 0) ++j;
    }
    res.push_back(res.front());
    return res;
}
void ring_halfplanes(const Ring& ring, double* A8, double* B4) {           // PCCBlackmoreSVC.cpp:77-108
    for (size_t i = 0; i + 1 < ring.size() && i < 4; ++i) {
        const double x1 = ring[i].x, y1 = ring[i].y, x2 = ring[i + 1].x, y2 = ring[i + 1].y;
        const double a = y2 - y1, b = x1 - x2;
        const double c = a * x1 + b * y1;
        A8[2 * i] = a; A8[2 * i + 1] = b; B4[i] = c;
    }
}
}  // namespace

extern "C" double cck_rect_robot_bounding_radius(double sx, double sy, double len, double width) {
    Ring shape = rect_ring(sx, sy, len, width);
    double max_rad = 0.0;
    for (const P2& p : shape) {                                            // common.cpp:68-85
        const double x = 1.0 * p.x + 0.0 * p.y + (0 - sx);
        const double y = -0.0 * p.x + 1.0 * p.y + (0 - sy);
        const double r = std::sqrt((x * x) + (y * y));
        if (r > max_rad) max_rad = r;
    }
    return max_rad;
}

extern "C" int cck_build_obstacle_tables(int M, const double* centres, double len, double width, double robot_rad, double* A,
                                         double* B, double* rects) {
    if (M < 0 || (M > 0 && (!centres || !A || !B))) return CCK_ERR_INVALID;
    const Ring robot = square_ring(robot_rad);
    for (int o = 0; o < M; ++o) {
        Ring m = minkowski(robot, rect_ring(centres[2 * o], centres[2 * o + 1], len, width));
        ring_halfplanes(m, A + 8 * o, B + 4 * o);
        if (rects) { rects[4 * o] = m[0].x; rects[4 * o + 1] = m[0].y; rects[4 * o + 2] = m[2].x; rects[4 * o + 3] = m[2].y; }
    }
    return CCK_OK;
}

extern "C" int cck_build_pair_halfplanes(int bounding_box, double r1, double r2, double* A8, double* B4) {
    if (!A8 || !B4) return CCK_ERR_INVALID;
    Ring poly = bounding_box ? square_ring(r1 + r2) : minkowski(square_ring(r1), square_ring(r2));
    ring_halfplanes(poly, A8, B4);
    return CCK_OK;
}

extern "C" int cck_model_params_default(int model, double dt, cck_model_params* p) {
    if (!p) return CCK_ERR_INVALID;
    std::memset(p, 0, sizeof(*p));
    p->model = model;
    p->dt = dt;
    const double q = std::pow(0.1, 2), r = std::pow(0.1, 2);
    if (model == CCK_MODEL_LINEAR2D) {
        p->dim = 2;
        const double K = 0.3;                                               // UncertainLinearSP.h:50
        p->lin_Q[0] = q * 1.0; p->lin_Q[3] = q * 1.0; p->lin_R[0] = r * 1.0; p->lin_R[3] = r * 1.0;
        p->lin_Acl[0] = 1.0 - 1.0 * K; p->lin_Acl[1] = 0.0 - 0.0 * K; p->lin_Acl[2] = 0.0 - 0.0 * K; p->lin_Acl[3] = 1.0 - 1.0 * K;
        return CCK_OK;
    }
    if (model == CCK_MODEL_UNICYCLE) {
        p->dim = 4;
        for (int i = 0; i < 4; ++i) p->uni_F[i * 4 + i] = 1.0;
        p->uni_F[1] = dt; p->uni_F[2 * 4 + 3] = dt;                         // exp(A_ol*dt), A_ol nilpotent (:88-91)
        const double Bd[4][2] = {{dt * dt / 2.0, 0.0}, {dt, 0.0}, {0.0, dt * dt / 2.0}, {0.0, dt}};   // :95-104
        const double Kg[2][4] = {{2.5857, 3.4434, 0.0, 0.0}, {0.0, 0.0, 2.5857, 3.4434}};             // :58-65
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) p->uni_Acl[i * 4 + j] = p->uni_F[i * 4 + j] - (Bd[i][0] * Kg[0][j] + Bd[i][1] * Kg[1][j]);
        for (int i = 0; i < 4; ++i) { p->uni_Q[i * 4 + i] = q * 1.0; p->uni_R[i * 4 + i] = r * 1.0; }
        const double g[8] = {1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0, 1.0};       // :76-83
        for (int i = 0; i < 8; ++i) p->uni_gains[i] = g[i];
        p->uni_acc[0] = -0.5; p->uni_acc[1] = 0.5; p->uni_turn[0] = -0.5; p->uni_turn[1] = 0.5;
        return CCK_OK;
    }
    return CCK_ERR_INVALID;
}

extern "C" int cck_set_model(cck_ctx* c, const cck_model_params* p) {
    if (!c || !p) return CCK_ERR_INVALID;
    if (p->model == CCK_MODEL_LINEAR2D && p->dim != 2) return fail(c, CCK_ERR_INVALID, "linear model needs dim 2");
    if (p->model == CCK_MODEL_UNICYCLE && p->dim != 4) return fail(c, CCK_ERR_INVALID, "unicycle model needs dim 4");
    if (p->model != CCK_MODEL_LINEAR2D && p->model != CCK_MODEL_UNICYCLE) return fail(c, CCK_ERR_INVALID, "unknown model");
    c->model = *p;
    c->lin.dt = p->dt;
    std::memcpy(c->lin.Q, p->lin_Q, sizeof(c->lin.Q)); std::memcpy(c->lin.R, p->lin_R, sizeof(c->lin.R));
    std::memcpy(c->lin.Acl, p->lin_Acl, sizeof(c->lin.Acl));
    {   // lin_step_diag: off-diagonals exactly +0.0 (bit pattern), |A_cl| <= 1; CCK_LIN_DENSE=1 forces the dense step
        const double z = 0.0;
        auto pz = [&](double v) { return std::memcmp(&v, &z, 8) == 0; };
        const double* q = c->lin.Q; const double* r = c->lin.R; const double* a = c->lin.Acl;
        const char* e = std::getenv("CCK_LIN_DENSE");
        c->lin.diag = (pz(q[1]) && pz(q[2]) && pz(r[1]) && pz(r[2]) && pz(a[1]) && pz(a[2]) && std::fabs(a[0]) <= 1.0 && std::fabs(a[3]) <= 1.0 &&
                       std::isfinite(q[0]) && std::isfinite(q[3]) && std::isfinite(r[0]) && std::isfinite(r[3]) && !(e && e[0] == '1')) ? 1 : 0;
        c->lin.pad = 0;
        c->lin.dg[0] = q[0]; c->lin.dg[1] = q[3]; c->lin.dg[2] = r[0]; c->lin.dg[3] = r[3]; c->lin.dg[4] = a[0]; c->lin.dg[5] = a[3];
    }
    std::memcpy(c->uni.F, p->uni_F, sizeof(c->uni.F)); std::memcpy(c->uni.Acl, p->uni_Acl, sizeof(c->uni.Acl));
    std::memcpy(c->uni.Q, p->uni_Q, sizeof(c->uni.Q)); std::memcpy(c->uni.R, p->uni_R, sizeof(c->uni.R));
    std::memcpy(c->uni.gains, p->uni_gains, sizeof(c->uni.gains));
    c->uni.acc_lo = p->uni_acc[0]; c->uni.acc_hi = p->uni_acc[1]; c->uni.turn_lo = p->uni_turn[0]; c->uni.turn_hi = p->uni_turn[1];
    // the sparse covariance step (cck_rollout_uni.cuh) needs the reference's exact-zero pattern
    {
        const double* F = c->uni.F; const double* A = c->uni.Acl;
        bool ok = F[0] == 1.0 && F[5] == 1.0 && F[10] == 1.0 && F[15] == 1.0;
        for (int i = 0; i < 16; ++i) {
            const int r = i / 4, cc = i % 4;
            const bool f_nz = (r == cc) || i == 1 || i == 11;
            const bool a_nz = (r / 2) == (cc / 2);
            if (!f_nz && F[i] != 0.0) ok = false;
            if (!a_nz && A[i] != 0.0) ok = false;
            if (r != cc && (c->uni.Q[i] != 0.0 || c->uni.R[i] != 0.0)) ok = false;
            if (!std::isfinite(F[i]) || !std::isfinite(A[i]) || !std::isfinite(c->uni.Q[i]) || !std::isfinite(c->uni.R[i])) ok = false;
        }
        c->uni_sparse = ok;
        c->uni.sparse = ok ? 1 : 0; c->uni.pad = 0;
    }
    c->has_model = true;
    if (c->d_cov_tab) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream); cudaFree(c->d_cov_tab); c->d_cov_tab = nullptr; c->cov_tab_n = 0; }   // built for the old model
    return CCK_OK;
}


// ---------------------------------------------------------------------------
// fp32 box + cell-grid table of the Blackmore broad phase (cck_rollout_grid.cuh).
// Box sides are rounded OUTWARD and bumped one more ulp, so "interval misses box" in fp32
// implies the reference's fp64 inequality for that half-plane.
// ---------------------------------------------------------------------------
static float f32_up(double v) {      // smallest-ish float strictly above v (>= v + 1 ulp)
    float f = (float)v;
    if ((double)f < v) f = std::nextafterf(f, HUGE_VALF);
    return std::nextafterf(f, HUGE_VALF);
}
static float f32_down(double v) {
    float f = (float)v;
    if ((double)f > v) f = std::nextafterf(f, -HUGE_VALF);
    return std::nextafterf(f, -HUGE_VALF);
}

struct GridBuild { std::vector<unsigned char> buf; std::vector<double> exB; std::vector<int32_t> exCls; std::vector<int32_t> order; };

static void build_grid_table(int M, const double* A, const double* B, const std::vector<int>& cls_idx, double c, GridBuild& out) {
    const bool c_ok = std::isfinite(c) && c >= 0.0 && c < 1e6;
    struct Box { float xlo, ylo, xhi, yhi; };
    std::vector<Box> box(M), inner(M);
    std::vector<int> gridded, always;
    auto usable = [](double w) { const double a = std::fabs(w); return std::isfinite(w) && a > 1e-100 && a < 1e100; };
    for (int o = 0; o < M; ++o) {
        Box b{-HUGE_VALF, -HUGE_VALF, HUGE_VALF, HUGE_VALF};
        Box q{HUGE_VALF, HUGE_VALF, -HUGE_VALF, -HUGE_VALF};   // inner box: "every plane fails"; stays empty unless one plane per side
        int sides = 0;                                          // bit per side that has exactly one usable plane
        bool pure = c_ok;
        for (int i = 0; i < 4 && c_ok; ++i) {
            const double a0 = A[8 * o + 2 * i], a1 = A[8 * o + 2 * i + 1], bv = B[4 * o + i];
            if (std::isnan(bv)) { pure = false; continue; }
            if (a1 == 0.0 && usable(a0)) {            // a0*x >= b + bb, bb/|a0| = sqrt2*c*sqrt(P00)
                const double thr = bv / a0;
                if (a0 > 0) { b.xhi = std::min(b.xhi, f32_up(thr)); q.xhi = f32_down(thr); pure = pure && !(sides & 1); sides |= 1; }
                else { b.xlo = std::max(b.xlo, f32_down(thr)); q.xlo = f32_up(thr); pure = pure && !(sides & 2); sides |= 2; }
            } else if (a0 == 0.0 && usable(a1)) {     // a1*y >= b + bb, bb/|a1| = sqrt2*c*sqrt(P11)
                const double thr = bv / a1;
                if (a1 > 0) { b.yhi = std::min(b.yhi, f32_up(thr)); q.yhi = f32_down(thr); pure = pure && !(sides & 4); sides |= 4; }
                else { b.ylo = std::max(b.ylo, f32_down(thr)); q.ylo = f32_up(thr); pure = pure && !(sides & 8); sides |= 8; }
            } else pure = false;
        }
        if (!pure || sides != 15) q = Box{HUGE_VALF, HUGE_VALF, -HUGE_VALF, -HUGE_VALF};
        // sides beyond the magnitudes the broad phase accepts (|x|, P < 1e30) become infinite, so that one far-away
        // plane cannot stretch the cell grid: towards "never proven" from 1e30 on (always allowed), towards "always
        // proven" only beyond 1e31 (x -+ ex stays below 1.01e30 for every belief that passes the finiteness guard)
        auto clamp_hi = [](float v) { return v > 1e30f ? HUGE_VALF : (v < -1e31f ? -HUGE_VALF : v); };   // proven if x - ex >= v
        auto clamp_lo = [](float v) { return v < -1e30f ? -HUGE_VALF : (v > 1e31f ? HUGE_VALF : v); };   // proven if x + ex <= v
        b.xhi = clamp_hi(b.xhi); b.yhi = clamp_hi(b.yhi); b.xlo = clamp_lo(b.xlo); b.ylo = clamp_lo(b.ylo);
        if (!(std::fabs(b.xlo) < 1e30f && std::fabs(b.ylo) < 1e30f && std::fabs(b.xhi) < 1e30f && std::fabs(b.yhi) < 1e30f))
            q = Box{HUGE_VALF, HUGE_VALF, -HUGE_VALF, -HUGE_VALF};
        box[o] = b; inner[o] = q;
        const bool finite = std::isfinite(b.xlo) && std::isfinite(b.ylo) && std::isfinite(b.xhi) && std::isfinite(b.yhi);
        (finite ? gridded : always).push_back(o);
    }
    GHeader h{};
    h.M = M; h.n_grid = (int)gridded.size(); h.R = CCK_GRID_ROWS;
    h.k = c_ok ? f32_up(1.4142135623730951 * c * (1.0 + 1e-15)) : 0.f;
    double sx = 0, sy = 0, x0 = HUGE_VAL, x1 = -HUGE_VAL, y0 = HUGE_VAL, y1 = -HUGE_VAL;
    for (int o : gridded) {
        sx = std::max(sx, (double)box[o].xhi - (double)box[o].xlo); sy = std::max(sy, (double)box[o].yhi - (double)box[o].ylo);
        x0 = std::min(x0, (double)box[o].xlo); x1 = std::max(x1, (double)box[o].xlo);
        y0 = std::min(y0, (double)box[o].ylo); y1 = std::max(y1, (double)box[o].ylo);
    }
    if (gridded.empty()) { x0 = x1 = y0 = y1 = 0; }
    h.SX = f32_up(sx); h.SY = f32_up(sy);
    h.gx0 = (float)x0; h.gy0 = (float)y0;
    if ((double)h.gx0 > x0) h.gx0 = std::nextafterf(h.gx0, -HUGE_VALF);
    if ((double)h.gy0 > y0) h.gy0 = std::nextafterf(h.gy0, -HUGE_VALF);
    // cell size.  Rows: the whole anchor range must fit 64 rows, and a row is not finer than the typical box (every grid
    // row a query overlaps costs a CSR lookup).  Columns: always all 64 over the anchor range - within a row the
    // candidates are ONE contiguous CSR range whatever the column width, so finer columns only trim false candidates.
    auto inv_cell = [](double range, double s, int n, bool clamp_to_box) {
        double ic = 1.0 / std::max(s, 1e-30);
        if (range > 0) ic = clamp_to_box ? std::min(ic, (n - 0.5) / range) : (n - 0.5) / range;
        return (float)std::min(ic, 1e30);
    };
    h.icx = inv_cell(x1 - x0, sx, CCK_GRID_COLS, false); h.icy = inv_cell(y1 - y0, sy, CCK_GRID_ROWS, true);
    auto fits = [&]() {
        for (int o : gridded) {
            const float cx = grid_cellf(box[o].xlo, h.gx0, h.icx), cy = grid_cellf(box[o].ylo, h.gy0, h.icy);
            if (!(cx >= 0.f && cx <= (float)(CCK_GRID_COLS - 1) && cy >= 0.f && cy <= (float)(CCK_GRID_ROWS - 1))) return false;
        }
        return true;
    };
    for (int it = 0; it < 64 && !fits(); ++it) { h.icx *= 0.999f; h.icy *= 0.999f; }
    if (!fits()) {   // cannot bin (pathological coordinates): every obstacle goes through the box test directly
        always.insert(always.end(), gridded.begin(), gridded.end());
        gridded.clear(); h.n_grid = 0;
    }
    struct Ent { int cy, cx, o; };
    std::vector<Ent> ents;
    for (int o : gridded) ents.push_back({(int)grid_cellf(box[o].ylo, h.gy0, h.icy), (int)grid_cellf(box[o].xlo, h.gx0, h.icx), o});
    std::stable_sort(ents.begin(), ents.end(), [](const Ent& a, const Ent& b) { return a.cy != b.cy ? a.cy < b.cy : a.cx < b.cx; });
    // dense CSR over the R x 64 cells in row-major order (ents are sorted the same way): the obstacles whose anchor lies in
    // row r, columns c0..c1 are the contiguous rows dstart[r*64 + c0] .. dstart[r*64 + c1 + 1] - 1 of the box arrays
    const int R = ents.empty() ? 0 : ents.back().cy + 1;
    std::vector<uint16_t> dstart((size_t)R * CCK_GRID_COLS + 1, 0);
    {
        size_t e = 0;
        for (int idx = 0; idx <= R * CCK_GRID_COLS; ++idx) {
            while (e < ents.size() && ents[e].cy * CCK_GRID_COLS + ents[e].cx < idx) ++e;
            dstart[idx] = (uint16_t)e;
        }
    }
    h.R = R;
    h.n_cells = R * CCK_GRID_COLS;
    h.Rm1 = (float)(R - 1);      // -1 for an empty grid: every window is empty
    auto al = [](int v, int a) { return (v + a - 1) / a * a; };
    h.off_occ = 0; h.off_rowbase = 0;
    h.off_cstart = (int)sizeof(GHeader);
    h.off_boxes = al(h.off_cstart + (int)dstart.size() * 2, 16);
    h.total_bytes = al(h.off_boxes + 2 * M * 16, 16);
    out.buf.assign(h.total_bytes, 0);
    std::memcpy(out.buf.data(), &h, sizeof(h));
    std::memcpy(out.buf.data() + h.off_cstart, dstart.data(), dstart.size() * 2);
    out.exB.assign((size_t)std::max(M, 1) * 4, 0.0);
    out.exCls.assign(std::max(M, 1), 0);
    int j = 0;
    auto put = [&](int o) {
        std::memcpy(out.buf.data() + h.off_boxes + (size_t)j * 16, &box[o], 16);
        std::memcpy(out.buf.data() + h.off_boxes + (size_t)(M + j) * 16, &inner[o], 16);
        std::memcpy(out.exB.data() + (size_t)j * 4, B + 4 * o, 32);
        out.exCls[j] = cls_idx[o];
        out.order.push_back(o);
        ++j;
    };
    for (auto& e : ents) put(e.o);
    for (int o : always) put(o);
}

// ---------------------------------------------------------------------------
// Clearance field: CCK_DF_N x CCK_DF_N cells over the fp32 (inward-rounded) state bounds.  For obstacle o (table row o) and a
// cell, D_o = max(dx_o, dy_o) >= 0 with dx_o / dy_o = the gap between the (padded) cell interval and the OUTER box interval
// per axis: a LOWER bound of the Chebyshev clearance between any mean the device maps to the cell and that box.  A belief
// with max(ex, ey) <= D_o (ex = sqrt2 c sqrt(P00), ...) has x - ex >= xhi or x + ex <= xlo or the same in y: exactly what
// the box test of obstacle o proves.  Cell word = floor(16 D1) | floor(16 D2) << 8 | j1 << 16 with D1 <= D2 the two smallest
// D_o and j1 the row attaining D1 (values saturate at 255; one obstacle: D2 = saturated).  The cell is padded by `pad` per side for the fp32
// roundings of the device's cell function and of the mean itself; the field is built only when every obstacle has a finite
// outer box (n_grid == M) and the bounds are usable.  Returns false when there is no field.
// ---------------------------------------------------------------------------
static bool build_clearance_field(const GridBuild& gb, double c, float blx, float bhx, float bly, float bhy, std::vector<uint32_t>& q,
                                  float& icx, float& icy, float& k2) {
    GHeader h;
    std::memcpy(&h, gb.buf.data(), sizeof(h));
    const int M = h.M, N = CCK_DF_N;
    const bool c_ok = std::isfinite(c) && c >= 0.0 && c < 1e6;
    if (M <= 0 || h.n_grid != M || !c_ok) return false;
    const double rx = (double)bhx - (double)blx, ry = (double)bhy - (double)bly;
    const double mag = std::max(std::max(std::fabs((double)blx), std::fabs((double)bhx)), std::max(std::fabs((double)bly), std::fabs((double)bhy)));
    if (!(rx > 0 && ry > 0) || !std::isfinite(rx) || !std::isfinite(ry) || !(mag < 1e6) || !(rx > 1e-3 && ry > 1e-3)) return false;
    // cells per unit, shrunk by 2^-18 and rounded down: (xf - blx) * icx < N for every float xf < bhx
    icx = f32_down((double)N / rx * (1.0 - 1.0 / 262144.0));
    icy = f32_down((double)N / ry * (1.0 - 1.0 / 262144.0));
    k2 = f32_up(CCK_DF_SCALE * CCK_DF_SCALE * 2.0 * c * c * (1.0 + 1.0 / 262144.0) * (1.0 + 1e-12));
    const double wx = 1.0 / (double)icx, wy = 1.0 / (double)icy;          // cell size in world units
    const double padx = wx / 1024.0 + (std::max(mag, 1.0) + rx) / 1048576.0, pady = wy / 1024.0 + (std::max(mag, 1.0) + ry) / 1048576.0;
    const float* boxes = reinterpret_cast<const float*>(gb.buf.data() + h.off_boxes);
    std::vector<float> D((size_t)N * N, HUGE_VALF), D2((size_t)N * N, HUGE_VALF);
    std::vector<int> J((size_t)N * N, 0);
    const double reach = 256.0 / CCK_DF_SCALE;      // clearances beyond 255/16 saturate
    for (int o = 0; o < M; ++o) {
        const double xlo = boxes[4 * o], ylo = boxes[4 * o + 1], xhi = boxes[4 * o + 2], yhi = boxes[4 * o + 3];
        // only cells within `reach` of the box can get a value below saturation from it
        const int i0 = std::max(0, (int)std::floor((xlo - reach - blx) / wx) - 1), i1 = std::min(N - 1, (int)std::floor((xhi + reach - blx) / wx) + 1);
        const int j0 = std::max(0, (int)std::floor((ylo - reach - bly) / wy) - 1), j1 = std::min(N - 1, (int)std::floor((yhi + reach - bly) / wy) + 1);
        for (int j = j0; j <= j1; ++j) {
            const double cy0 = (double)bly + j * wy - pady, cy1 = (double)bly + (j + 1) * wy + pady;
            const double dy = std::max(ylo - cy1, cy0 - yhi);
            for (int i = i0; i <= i1; ++i) {
                const double cx0 = (double)blx + i * wx - padx, cx1 = (double)blx + (i + 1) * wx + padx;
                const double dx = std::max(xlo - cx1, cx0 - xhi);
                const float d = (float)std::max(0.0, std::max(dx, dy));
                const size_t ci = (size_t)j * N + i;
                if (d < D[ci]) { D2[ci] = D[ci]; D[ci] = d; J[ci] = o; }
                else if (d < D2[ci]) D2[ci] = d;
            }
        }
    }
    q.assign((size_t)N * N, 0);
    auto quant = [&](float d) { return (uint32_t)std::max(0.0, std::min(255.0, std::floor(std::min((double)d, reach) * CCK_DF_SCALE * (1.0 - 1e-6)))); };
    for (size_t i = 0; i < q.size(); ++i) q[i] = quant(D[i]) | (quant(D2[i]) << 8) | ((uint32_t)J[i] << 16);
    return true;
}

// Host-only (no GPU, no context): the clearance field cck_set_obstacles builds next to the grid table, for CPU tests of its
// invariant (a cell's values never exceed 16 x the true Chebyshev clearances of any point mapped to it).  out: CCK_DF_N^2
// 32-bit words (row-major, y-major); params[0..2] = icx, icy, k2.  Returns CCK_OK with *has = 0 when no field is built.
extern "C" int cck_build_clearance_field(int M, const double* A, const double* B, double erf_inv_result, const double* low, const double* high,
                                         uint32_t* out, float* params, int* has) {
    if (M < 0 || (M > 0 && (!A || !B)) || !low || !high || !out || !params || !has) return CCK_ERR_INVALID;
    *has = 0;
    if (M == 0) return CCK_OK;
    GridBuild gb;
    std::vector<int> cls_idx(M, 0);
    build_grid_table(M, A, B, cls_idx, erf_inv_result, gb);
    std::vector<uint32_t> q;
    float icx, icy, k2;
    if (!build_clearance_field(gb, erf_inv_result, f32_up(low[0]), f32_down(high[0]), f32_up(low[1]), f32_down(high[1]), q, icx, icy, k2)) return CCK_OK;
    std::memcpy(out, q.data(), q.size() * 4);
    params[0] = icx; params[1] = icy; params[2] = k2;
    *has = 1;
    return CCK_OK;
}

// Host-only: the serialised grid table cck_set_obstacles would upload (no GPU, no context needed), so that its
// invariants can be tested on a CPU-only machine: layout = GHeader (80 B: int32 total_bytes, M, n_grid, n_cells = R*64, R,
// 0, 0, off_cstart, off_boxes, pad; float gx0, gy0, icx, icy, SX, SY, k, Rm1; 2 x int32 pad), uint16 dstart[R*64+1] (dense
// row-major CSR over the cells), float4 outer[M], float4 inner[M] (xlo, ylo, xhi, yhi).  order[j] = index of the obstacle
// stored in row j.
extern "C" int cck_build_grid_table(int M, const double* A, const double* B, double erf_inv_result, unsigned char* out, int64_t cap,
                                    int64_t* bytes, int32_t* order) {
    if (M < 0 || (M > 0 && (!A || !B)) || !bytes) return CCK_ERR_INVALID;
    GridBuild gb;
    std::vector<int> cls_idx(std::max(M, 1), 0);
    build_grid_table(M, A, B, cls_idx, erf_inv_result, gb);
    *bytes = (int64_t)gb.buf.size();
    if (out) {
        if (cap < (int64_t)gb.buf.size()) return CCK_ERR_INVALID;
        std::memcpy(out, gb.buf.data(), gb.buf.size());
    }
    if (order) for (int j = 0; j < M; ++j) order[j] = gb.order[j];
    return CCK_OK;
}

// Group obstacles by bit-identical A[4][2] and lay the table out for one TMA bulk copy.
extern "C" int cck_set_obstacles(cck_ctx* c, const cck_svc_params* p, const double* A, const double* B, const double* centres,
                                 const double* rects, int M) {
    if (!c || !p || M < 0) return CCK_ERR_INVALID;
    if (M > 0 && (!A || !B)) return fail(c, CCK_ERR_INVALID, "A/B tables missing");
    if (p->kind == CCK_SVC_ADAPTIVE && M > 0 && !centres) return fail(c, CCK_ERR_INVALID, "Adaptive SVC needs obstacle centres");
    if (p->kind == CCK_SVC_CHISQUARED && M > 0 && !rects) return fail(c, CCK_ERR_INVALID, "ChiSquared SVC needs inflated rectangles");
    if (p->kind < 0 || p->kind > 2 || p->dim < 1 || p->dim > 4) return fail(c, CCK_ERR_INVALID, "bad svc params");
    if (M > 65535) return fail(c, CCK_ERR_INVALID, "more than 65535 obstacles");   // uint16 cell CSR of the grid table
    CK(c, cudaSetDevice(c->device));

    struct Key { uint64_t w[8]; bool operator<(const Key& o) const { return std::memcmp(w, o.w, sizeof(w)) < 0; } };
    std::map<Key, int> cls_of;
    std::vector<std::vector<int>> members;
    std::vector<Key> keys;
    std::vector<int> cls_idx(M);
    for (int o = 0; o < M; ++o) {
        Key k;
        std::memcpy(k.w, A + 8 * o, 64);
        auto it = cls_of.find(k);
        if (it == cls_of.end()) { it = cls_of.emplace(k, (int)members.size()).first; members.emplace_back(); keys.push_back(k); }
        members[it->second].push_back(o);
        cls_idx[o] = it->second;
    }
    const int nc = (int)members.size();
    int rows = 0;   // B rows incl. per-class padding to a multiple of 4
    for (int k = 0; k < nc; ++k) rows += ((int)members[k].size() + 3) & ~3;
    TabHeader h{};
    h.n_classes = nc; h.M = M;
    h.off_classes = sizeof(TabHeader);
    h.off_B = h.off_classes + nc * (int)sizeof(ClassRec);
    h.off_centres = h.off_B + rows * 32;
    h.off_rects = h.off_centres + (centres ? M * 16 : 0);
    h.total_bytes = h.off_rects + (rects ? M * 32 : 0);
    h.total_bytes = (h.total_bytes + 15) & ~15;
    h.pad = rows;
    // the class table is staged into shared memory only by the generic kernels (checked at their launch);
    // the default Blackmore rollout reads its class normals from global memory on the exact path
    std::vector<unsigned char> buf(h.total_bytes, 0);
    std::memcpy(buf.data(), &h, sizeof(h));
    int pos = 0;
    const double ninf[4] = {-HUGE_VAL, -HUGE_VAL, -HUGE_VAL, -HUGE_VAL};
    for (int k = 0; k < nc; ++k) {
        ClassRec r{};
        std::memcpy(r.A, keys[k].w, 64);
        r.first = pos; r.count = (int)members[k].size(); r.count_pad = (r.count + 3) & ~3;
        std::memcpy(buf.data() + h.off_classes + k * sizeof(ClassRec), &r, sizeof(r));
        for (int o : members[k]) { std::memcpy(buf.data() + h.off_B + pos * 32, B + 4 * o, 32); ++pos; }
        for (int j = r.count; j < r.count_pad; ++j) { std::memcpy(buf.data() + h.off_B + pos * 32, ninf, 32); ++pos; }
    }
    if (centres && M) std::memcpy(buf.data() + h.off_centres, centres, (size_t)M * 16);   // original order (sum order, eta_list[0])
    if (rects && M) std::memcpy(buf.data() + h.off_rects, rects, (size_t)M * 32);

    // the grid table is built (host only) and checked against the shared-memory budget BEFORE any context state changes:
    // a refused call leaves the previous tables installed and usable
    GridBuild gb;
    GHeader gh0{};
    bool boxes_in_smem = true;
    int stage_bytes = 0;
    if (M > 0) {
        build_grid_table(M, A, B, cls_idx, p->erf_inv_result, gb);
        std::memcpy(&gh0, gb.buf.data(), sizeof(gh0));
        // small tables are staged whole; for large M only the header + cell CSR (<= 8.3 KB) go to shared memory
        // and the box arrays (32 B per obstacle) are read from global memory
        boxes_in_smem = gb.buf.size() <= 24 * 1024;
        stage_bytes = boxes_in_smem ? (int)gb.buf.size() : gh0.off_boxes;
        if (grid_smem_bytes(stage_bytes, CCK_GRID_THREADS) > c->max_smem_optin) return fail(c, CCK_ERR_INVALID, "grid table exceeds shared memory");
    }
    CK(c, cudaStreamSynchronize(c->stream));
    c->has_svc = false;   // until every table below is in place: a CUDA failure half-way must not leave a usable-looking checker
    cudaFree(c->d_tab); c->d_tab = nullptr;
    CK(c, cudaMalloc(&c->d_tab, h.total_bytes));
    CK(c, cudaMemcpy(c->d_tab, buf.data(), h.total_bytes, cudaMemcpyHostToDevice));
    c->tab_bytes = h.total_bytes;
    c->M = M;

    // ---- fp32 box + cell-grid table (the default broad phase) ----
    cudaFree(c->d_df); c->d_df = nullptr;
    cudaFree(c->d_gtab); cudaFree(c->d_gexB); cudaFree(c->d_gexCls);
    c->d_gtab = nullptr; c->d_gexB = nullptr; c->d_gexCls = nullptr; c->gtab_bytes = 0;
    c->gconst = GridConst{};
    c->gconst.Rm1 = -1.f;   // no grid rows: every query window is empty (M = 0 keeps the one-branch fast exit)
    c->gconst.blx = f32_up(p->low[0]); c->gconst.bhx = f32_down(p->high[0]); c->gconst.bly = f32_up(p->low[1]); c->gconst.bhy = f32_down(p->high[1]);
    if (M > 0) {
        CK(c, cudaMalloc(&c->d_gtab, gb.buf.size()));
        CK(c, cudaMalloc(&c->d_gexB, gb.exB.size() * 8));
        CK(c, cudaMalloc(&c->d_gexCls, gb.exCls.size() * 4));
        CK(c, cudaMemcpy(c->d_gtab, gb.buf.data(), gb.buf.size(), cudaMemcpyHostToDevice));
        CK(c, cudaMemcpy(c->d_gexB, gb.exB.data(), gb.exB.size() * 8, cudaMemcpyHostToDevice));
        CK(c, cudaMemcpy(c->d_gexCls, gb.exCls.data(), gb.exCls.size() * 4, cudaMemcpyHostToDevice));
        c->gtab_bytes = stage_bytes;
        const GHeader& gh = gh0;
        const GridConst keep = c->gconst;
        c->gconst = GridConst{gh.k, gh.SX, gh.SY, gh.gx0, gh.gy0, gh.icx, gh.icy, gh.Rm1, gh.M, gh.n_grid, gh.off_occ, gh.off_rowbase, gh.off_cstart, gh.off_boxes,
                              keep.blx, keep.bhx, keep.bly, keep.bhy, (int32_t)sizeof(TabHeader), gh.Rm1 + 1.f,
                              boxes_in_smem ? nullptr : reinterpret_cast<const float4*>(c->d_gtab + gh.off_boxes),
                              nullptr, 0.f, 0.f, 0.f};
        std::vector<uint32_t> dfq;
        float dicx = 0, dicy = 0, dk2 = 0;
        if (!std::getenv("CCK_NO_DF") && build_clearance_field(gb, p->erf_inv_result, keep.blx, keep.bhx, keep.bly, keep.bhy, dfq, dicx, dicy, dk2)) {
            CK(c, cudaMalloc(&c->d_df, dfq.size() * 4));
            CK(c, cudaMemcpy(c->d_df, dfq.data(), dfq.size() * 4, cudaMemcpyHostToDevice));
            c->gconst.df = c->d_df; c->gconst.df_icx = dicx; c->gconst.df_icy = dicy; c->gconst.df_k2 = dk2;
        }
    }

    SvcDev& s = c->svc;
    std::memset(&s, 0, sizeof(s));
    s.kind = p->kind; s.dim = p->dim;
    for (int i = 0; i < 4; ++i) { s.low[i] = p->low[i]; s.high[i] = p->high[i]; }
    s.erf_inv_result = p->erf_inv_result; s.p_collision = p->p_collision; s.sc = p->sc; s.max_rad = p->max_rad;
    // bg::strategy::buffer::point_circle(10): a = 0, a -= 2pi/10 (ChiSquaredBoundarySVC.cpp:58-62)
    const double two_pi = 2.0 * M_PI, diff = two_pi / 10.0;
    double a = 0;
    for (int i = 0; i < 10; ++i, a -= diff) { s.dec_cos[i] = std::cos(a); s.dec_sin[i] = std::sin(a); }
    c->has_svc = true;
    return CCK_OK;
}

extern "C" int cck_set_pvc(cck_ctx* c, const cck_pvc_params* p) {
    if (!c || !p) return CCK_ERR_INVALID;
    if (p->k < 1 || p->k > CCK_MAX_ROBOTS) return fail(c, CCK_ERR_INVALID, "k out of range");
    if (p->dim != 2 && p->dim != 4) return fail(c, CCK_ERR_INVALID, "pvc dim must be 2 or 4");
    if (p->kind < 0 || p->kind > 6) return fail(c, CCK_ERR_INVALID, "unknown pvc kind");
    if (p->kind == CCK_PVC_CDFGRID && (p->grid_n < 2 || p->grid_n > 64)) return fail(c, CCK_ERR_INVALID, "CDFGrid needs 2 <= grid_n <= 64");
    if (!p->radii || !p->map_order || (p->kind != CCK_PVC_CHISQUARED && p->kind != CCK_PVC_CDFGRID && (!p->pair_A || !p->pair_B)))
        return fail(c, CCK_ERR_INVALID, "pvc tables missing");
    CK(c, cudaSetDevice(c->device));
    const int k = p->k;
    c->pvc_host = *p;
    c->pvc_radii.assign(p->radii, p->radii + k);
    c->pvc_order.assign(p->map_order, p->map_order + k);
    c->pvc_A.assign((size_t)k * k * 8, 0.0); c->pvc_B.assign((size_t)k * k * 4, 0.0);
    if (p->pair_A) std::memcpy(c->pvc_A.data(), p->pair_A, c->pvc_A.size() * 8);
    if (p->pair_B) std::memcpy(c->pvc_B.data(), p->pair_B, c->pvc_B.size() * 8);
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_pvc_A); cudaFree(c->d_pvc_B); cudaFree(c->d_pvc_radii); cudaFree(c->d_pvc_order);
    c->d_pvc_A = c->d_pvc_B = c->d_pvc_radii = nullptr; c->d_pvc_order = nullptr;
    CK(c, cudaMalloc(&c->d_pvc_A, c->pvc_A.size() * 8)); CK(c, cudaMalloc(&c->d_pvc_B, c->pvc_B.size() * 8));
    CK(c, cudaMalloc(&c->d_pvc_radii, k * 8)); CK(c, cudaMalloc(&c->d_pvc_order, k * 4));
    CK(c, cudaMemcpy(c->d_pvc_A, c->pvc_A.data(), c->pvc_A.size() * 8, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->d_pvc_B, c->pvc_B.data(), c->pvc_B.size() * 8, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->d_pvc_radii, c->pvc_radii.data(), k * 8, cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->d_pvc_order, c->pvc_order.data(), k * 4, cudaMemcpyHostToDevice));
    PvcDev& d = c->pvc;
    d.kind = p->kind; d.k = k; d.dim = p->dim; d.n_pairs = k * (k - 1) / 2; d.grid_n = p->grid_n; d.pad = 0;
    d.c_pair = p->c_pair; d.p_coll_agnts = p->p_coll_agnts; d.sc = p->sc;
    d.radii = c->d_pvc_radii; d.pair_A = c->d_pvc_A; d.pair_B = c->d_pvc_B; d.map_order = c->d_pvc_order;
    c->has_pvc = true;
    // a new PVC invalidates the constraint table
    cudaFree(c->d_cons); c->d_cons = nullptr; c->cons_cap = 0;
    return CCK_OK;
}

// ---------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------
static int grid_for(const cck_ctx* c, int64_t n, int threads, int ctas_per_sm) {
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)c->num_sms * ctas_per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}
static cudaStream_t pick(cck_ctx* c, void* stream) { return stream ? (cudaStream_t)stream : c->stream; }

template <typename K>
static int set_smem(cck_ctx* c, K kernel, int bytes) {
    // The opt-in limit is a per-function attribute shared by every thread of the process, and contexts of different robots
    // launch the same kernels concurrently with different sizes: always raise it to the device maximum, never to this
    // call's size (a smaller value set by another thread would fail this thread's launch).
    if (bytes > c->max_smem_optin) return fail(c, CCK_ERR_INVALID, "kernel needs more shared memory than the device offers");
    if (bytes > 48 * 1024) CK(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin));
    return CCK_OK;
}

static int check_tab_fits(cck_ctx* c) {
    if (16 + c->tab_bytes > c->max_smem_optin)
        return fail(c, CCK_ERR_INVALID, "obstacle table (" + std::to_string(c->tab_bytes) + " B) exceeds shared memory; M too large for this checker");
    return CCK_OK;
}

static int check_ready(cck_ctx* c, bool need_svc) {
    if (!c) return CCK_ERR_INVALID;
    if (!c->has_model) return fail(c, CCK_ERR_INVALID, "cck_set_model has not been called");
    if (need_svc && !c->has_svc) return fail(c, CCK_ERR_INVALID, "cck_set_obstacles has not been called");
    if (need_svc && c->svc.dim != c->model.dim) return fail(c, CCK_ERR_INVALID, "svc dim != model dim");
    return CCK_OK;
}

// ---------------------------------------------------------------------------
// propagate
// ---------------------------------------------------------------------------
extern "C" int cck_propagate_dev(cck_ctx* c, void* stream, int64_t n, const double* bel, int64_t ld, const double* ctrl, int64_t ldc,
                                 double duration, double* out, int64_t ldo) {
    int rc = check_ready(c, false);
    if (rc) return rc;
    if (n < 0 || ld < n || ldc < n || ldo < n) return fail(c, CCK_ERR_INVALID, "bad n/ld");
    if (n == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = pick(c, stream);
    if (c->model.model == CCK_MODEL_LINEAR2D) {
        k_propagate_lin<<<grid_for(c, n, 256, 8), 256, 0, s>>>(n, bel, ld, ctrl, ldc, out, ldo, c->lin);
    } else {
        k_propagate_uni<<<grid_for(c, n, 128, 8), 128, 0, s>>>(n, bel, ld, ctrl, ldc, duration, out, ldo, c->uni);
    }
    c->launches++;
    CK(c, cudaGetLastError());
    return CCK_OK;
}

// ---------------------------------------------------------------------------
// isValid
// ---------------------------------------------------------------------------
extern "C" int cck_is_valid_dev(cck_ctx* c, void* stream, int64_t n, const double* bel, int64_t ld, uint8_t* valid) {
    int rc = check_ready(c, true);
    if (rc) return rc;
    if (n < 0 || ld < n) return fail(c, CCK_ERR_INVALID, "bad n/ld");
    if (n == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = pick(c, stream);
    const int dim = c->model.dim;
    if (c->svc.kind == CCK_SVC_BLACKMORE) {   // fp32 box/cell-grid broad phase + exact decision
        const int gsm = 16 + c->gtab_bytes;
        if ((rc = set_smem(c, k_is_valid_grid, gsm))) return rc;
        k_is_valid_grid<<<grid_for(c, n, 256, 4), 256, gsm, s>>>(n, dim, bel, ld, valid, c->d_gtab, c->gtab_bytes, c->d_tab, c->d_gexB, c->d_gexCls, c->svc, c->gconst);
        c->launches++;
        CK(c, cudaGetLastError());
        return CCK_OK;
    }
    if ((rc = check_tab_fits(c))) return rc;
    const int smem = 16 + c->tab_bytes;
    const int g = grid_for(c, n, 256, 4);
#define LAUNCH_VALID(KIND)                                                                        \
    do {                                                                                          \
        if ((rc = set_smem(c, k_is_valid<KIND>, smem))) return rc;                                \
        k_is_valid<KIND><<<g, 256, smem, s>>>(n, dim, bel, ld, valid, c->d_tab, c->tab_bytes, c->svc); \
    } while (0)
    if (c->svc.kind == CCK_SVC_ADAPTIVE) LAUNCH_VALID(CCK_SVC_ADAPTIVE);
    else LAUNCH_VALID(CCK_SVC_CHISQUARED);
#undef LAUNCH_VALID
    c->launches++;
    CK(c, cudaGetLastError());
    return CCK_OK;
}

// ---------------------------------------------------------------------------
// propagateWhileValid / batched expansion
// ---------------------------------------------------------------------------
template <bool EXPAND>
static int launch_rollout(cck_ctx* c, cudaStream_t s, const RolloutArgs& a) {
    int rc;
    const int smem = 16 + c->tab_bytes;
    const int threads = 128;
    const int g = grid_for(c, a.n, threads, 8);
    const bool traj = a.traj != nullptr;
#define LAUNCH_R(KERN, PAR)                                                         \
    do {                                                                            \
        if ((rc = set_smem(c, KERN, smem))) return rc;                              \
        KERN<<<g, threads, smem, s>>>(a, PAR, c->svc);                              \
    } while (0)
#define DISPATCH_KIND(NAME, PAR, TRAJ)                                                                              \
    do {                                                                                                            \
        if (c->svc.kind == CCK_SVC_ADAPTIVE) LAUNCH_R((NAME<CCK_SVC_ADAPTIVE, TRAJ, EXPAND>), PAR);           \
        else LAUNCH_R((NAME<CCK_SVC_CHISQUARED, TRAJ, EXPAND>), PAR);                                               \
    } while (0)
    if (c->model.model == CCK_MODEL_LINEAR2D && c->svc.kind == CCK_SVC_BLACKMORE) {
        // default configuration: fp32 box/cell-grid broad phase + exact fp64 decision, lane refill
        if (a.n > (int64_t)INT_MAX) return fail(c, CCK_ERR_INVALID, "more than 2^31-1 beliefs in one launch");
        RolloutArgs g2 = a;
        g2.ftab = c->d_gtab; g2.ftab_bytes = c->gtab_bytes; g2.exB = c->d_gexB; g2.exCls = c->d_gexCls;
        const int rl2 = rollout_range_log2(a.n);
        const int64_t nr = (a.n + ((int64_t)1 << rl2) - 1) >> rl2;
        const int gb = (int)std::min<int64_t>(nr, (int64_t)c->num_sms * grid_minb(EXPAND));
        const int gsmem = grid_smem_bytes(c->gtab_bytes, grid_threads(EXPAND)) + (EXPAND ? a.cons_smem_bytes : 0);
        if (traj) { if ((rc = set_smem(c, (k_rollout_lin_grid<true, EXPAND>), gsmem))) return rc;
                    k_rollout_lin_grid<true, EXPAND><<<gb, grid_threads(EXPAND), gsmem, s>>>(g2, c->lin, c->svc, c->gconst, rl2); }
        else { if ((rc = set_smem(c, (k_rollout_lin_grid<false, EXPAND>), gsmem))) return rc;
               k_rollout_lin_grid<false, EXPAND><<<gb, grid_threads(EXPAND), gsmem, s>>>(g2, c->lin, c->svc, c->gconst, rl2); }
    } else if (c->model.model == CCK_MODEL_LINEAR2D) {
        // adaptive / chi-squared checkers: generic kernel over the staged class table
        if ((rc = check_tab_fits(c))) return rc;
        if (traj) DISPATCH_KIND(k_rollout_lin, c->lin, true); else DISPATCH_KIND(k_rollout_lin, c->lin, false);
    } else {
        // unicycle: sparse covariance step where the model has the reference's exact-zero pattern (UniDev::sparse, else the
        // dense step), grid broad phase for Blackmore, lane refill
        RolloutArgs g2 = a;
        if (c->svc.kind == CCK_SVC_BLACKMORE) { g2.ftab = c->d_gtab; g2.ftab_bytes = c->gtab_bytes; g2.exB = c->d_gexB; g2.exCls = c->d_gexCls; }
        else { if ((rc = check_tab_fits(c))) return rc; g2.ftab = c->d_tab; g2.ftab_bytes = c->tab_bytes; }
        const int rl2 = rollout_range_log2(a.n);
        const int64_t nr = (a.n + ((int64_t)1 << rl2) - 1) >> rl2;
        const int gb = (int)std::min<int64_t>(nr, (int64_t)c->num_sms * 2);
        const int usmem = uni_smem_bytes(g2.ftab_bytes) + (EXPAND ? a.cons_smem_bytes : 0);
        if (usmem > c->max_smem_optin) return fail(c, CCK_ERR_INVALID, "obstacle table exceeds shared memory; M too large");
#define LAUNCH_U2(KIND, TRAJ)                                                                        \
    do {                                                                                             \
        if ((rc = set_smem(c, (k_rollout_uni2<KIND, TRAJ, EXPAND>), usmem))) return rc;              \
        k_rollout_uni2<KIND, TRAJ, EXPAND><<<gb, CCK_UNI_THREADS, usmem, s>>>(g2, c->uni, c->svc, c->gconst, rl2);   \
    } while (0)
        if (c->svc.kind == CCK_SVC_BLACKMORE) { if (traj) LAUNCH_U2(CCK_SVC_BLACKMORE, true); else LAUNCH_U2(CCK_SVC_BLACKMORE, false); }
        else if (c->svc.kind == CCK_SVC_ADAPTIVE) { if (traj) LAUNCH_U2(CCK_SVC_ADAPTIVE, true); else LAUNCH_U2(CCK_SVC_ADAPTIVE, false); }
        else { if (traj) LAUNCH_U2(CCK_SVC_CHISQUARED, true); else LAUNCH_U2(CCK_SVC_CHISQUARED, false); }
#undef LAUNCH_U2
    }
#undef DISPATCH_KIND
#undef LAUNCH_R
    c->launches++;
    CK(c, cudaGetLastError());
    return CCK_OK;
}

extern "C" int cck_propagate_while_valid_dev(cck_ctx* c, void* stream, int64_t n, const double* bel, int64_t ld, const double* ctrl,
                                             int64_t ldc, const int32_t* steps, double* out, int64_t ldo, int32_t* valid_steps,
                                             double* traj, int64_t traj_ld, int32_t traj_T) {
    int rc = check_ready(c, true);
    if (rc) return rc;
    if (n < 0 || ld < n || ldc < n || ldo < n || (traj && traj_ld < n)) return fail(c, CCK_ERR_INVALID, "bad n/ld");
    if (traj && traj_T < 1) return fail(c, CCK_ERR_INVALID, "traj_T must be >= 1 when a trajectory buffer is given");
    if (n == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    RolloutArgs a{};
    a.n = n; a.bel = bel; a.ld = ld; a.ctrl = ctrl; a.ldc = ldc; a.steps = steps; a.out = out; a.ldo = ldo; a.valid = valid_steps;
    a.traj = traj; a.traj_ld = traj_ld; a.traj_T = traj_T; a.tab = c->d_tab; a.tab_bytes = c->tab_bytes; a.duration = c->model.dt;
    return launch_rollout<false>(c, pick(c, stream), a);
}

extern "C" int cck_expand_batch_dev(cck_ctx* c, void* stream, int64_t n, const double* bel, int64_t ld, const double* ctrl, int64_t ldc,
                                    const int32_t* steps, const int32_t* parent_step, const double* parent_cost,
                                    const cck_goal_params* goal, double* out, int64_t ldo, int32_t* valid_steps, double* cost,
                                    double* goal_dist, uint8_t* flags) {
    int rc = check_ready(c, true);
    if (rc) return rc;
    if (!goal || !parent_step || !parent_cost || !cost || !goal_dist || !flags) return fail(c, CCK_ERR_INVALID, "null argument");
    if (n < 0 || ld < n || ldc < n || ldo < n) return fail(c, CCK_ERR_INVALID, "bad n/ld");
    if (n == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    RolloutArgs a{};
    a.n = n; a.bel = bel; a.ld = ld; a.ctrl = ctrl; a.ldc = ldc; a.steps = steps; a.out = out; a.ldo = ldo; a.valid = valid_steps;
    a.tab = c->d_tab; a.tab_bytes = c->tab_bytes; a.duration = c->model.dt;
    a.parent_step = parent_step; a.parent_cost = parent_cost; a.cost = cost; a.goal_dist = goal_dist; a.flags = flags;
    a.goal = *goal; a.cons = c->d_cons;
    a.cov_tab = c->d_cov_tab; a.cov_tab_n = c->cov_tab_n;
    a.cons_smem_bytes = (c->d_cons && c->cons_bytes <= CCK_CONS_SMEM_MAX) ? c->cons_bytes : 0;   // small tables are copied into shared memory by each CTA
    return launch_rollout<true>(c, pick(c, stream), a);
}

// Host-only test hook (no GPU, no context): the pair test of plan-validity-checker `kind`, evaluated by the SAME source
// functions the validatePlan / satisfiesConstraints kernels call (cck_kernels.cuh, compiled here for the host), so that their
// logic can be checked against the oracle on a CPU-only machine.  Kinds with a data-dependent risk allocation (adaptive) take
// c_pair as given.  Returns 1 safe, 0 unsafe, -1 bad kind.
extern "C" int cck_host_pair_safe(int kind, const double* mu_a, const double* cov_a, const double* mu_b, const double* cov_b, const double* A8,
                                  const double* B4, double rad_a, double rad_b, double c_pair, double p_coll_agnts, double sc, int grid_n) {
    const double dx = mu_a[0] - mu_b[0], dy = mu_a[1] - mu_b[1];
    const double S0 = cov_a[0] + cov_b[0], S1 = cov_a[1] + cov_b[1], S2 = cov_a[2] + cov_b[2], S3 = cov_a[3] + cov_b[3];
    switch (kind) {
        case CCK_PVC_BLACKMORE: case CCK_PVC_BOUNDINGBOX: case CCK_PVC_ADAPTIVE_BLACKMORE: case CCK_PVC_ADAPTIVE_BOUNDINGBOX:
            return pair_safe_halfplanes(A8, B4, dx, dy, S0, S1, S2, S3, c_pair) ? 1 : 0;
        case CCK_PVC_CHISQUARED: {
            const double la = max_eig_real_2x2(cov_a[0], cov_a[1], cov_a[2], cov_a[3]), lb = max_eig_real_2x2(cov_b[0], cov_b[1], cov_b[2], cov_b[3]);
            const double dist = std::sqrt(dx * dx + dy * dy);
            return dist > (la * sc + rad_a) + (lb * sc + rad_b) ? 1 : 0;
        }
        case CCK_PVC_BLACKMORE2: return pair_safe_blackmore2(A8, B4, dx, dy, S0, S1, S2, S3, p_coll_agnts) ? 1 : 0;
        case CCK_PVC_CDFGRID: return pair_safe_cdfgrid(mu_a, cov_a, mu_b, cov_b, rad_a + rad_b, grid_n, p_coll_agnts) ? 1 : 0;
        default: return -1;
    }
}

// Host-only test hook: CentralizedChiSquaredBoundarySVC's agent-agent decagon test by the kernel's own source (k_chisq_pairs)
extern "C" int cck_host_chisq_pair_disjoint(const double* mu_a, const double* cov_a, double rad_a, const double* mu_b, const double* cov_b, double rad_b, double sc) {
    double dc[10], ds[10];
    const double two_pi = 2.0 * M_PI, diff = two_pi / 10.0;
    double ang = 0;
    for (int i = 0; i < 10; ++i, ang -= diff) { dc[i] = std::cos(ang); ds[i] = std::sin(ang); }
    return chisq_pair_disjoint(mu_a[0], mu_a[1], cov_a[0], cov_a[1], cov_a[2], cov_a[3], rad_a, mu_b[0], mu_b[1], cov_b[0], cov_b[1], cov_b[2], cov_b[3], rad_b, sc, dc, ds) ? 1 : 0;
}

// Covariance table of a low-level tree (see RolloutArgs::cov_tab): entries for absolute steps 0 .. n_steps-1 from the root
// covariance root_cov = [Sigma row-major, Lambda row-major] (2 dim^2 doubles).  n_steps <= 0 removes the table.
extern "C" int cck_set_cov_table(cck_ctx* c, const double* root_cov, int n_steps) {
    int rc = check_ready(c, false);
    if (rc) return rc;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->d_cov_tab); c->d_cov_tab = nullptr; c->cov_tab_n = 0;
    if (n_steps <= 0) return CCK_OK;
    if (!root_cov || n_steps > (1 << 20)) return fail(c, CCK_ERR_INVALID, "bad covariance table arguments");
    const int dim = c->model.dim, per = 2 * dim * dim;
    double* d_root = nullptr;
    CK(c, cudaMalloc(&c->d_cov_tab, (size_t)n_steps * per * 8));
    CK(c, cudaMalloc(&d_root, per * 8));
    CK(c, cudaMemcpy(d_root, root_cov, per * 8, cudaMemcpyHostToDevice));
    if (c->model.model == CCK_MODEL_LINEAR2D) k_cov_table_lin<<<1, 32, 0, c->stream>>>(c->d_cov_tab, n_steps, c->lin, d_root);
    else k_cov_table_uni<<<1, 32, 0, c->stream>>>(c->d_cov_tab, n_steps, c->uni, d_root);
    c->launches++;
    const cudaError_t e = cudaStreamSynchronize(c->stream);
    cudaFree(d_root);
    if (e != cudaSuccess) { cudaFree(c->d_cov_tab); c->d_cov_tab = nullptr; c->err = std::string("k_cov_table: ") + cudaGetErrorString(e); return CCK_ERR_CUDA; }
    c->cov_tab_n = n_steps;
    return CCK_OK;
}

// ---------------------------------------------------------------------------
// validatePlan
// ---------------------------------------------------------------------------
// Scans the steps [t_begin, t_end) for the first conflict; the follow-up of the winning pair always covers the whole plan.
extern "C" int cck_validate_plan_range_dev(cck_ctx* c, void* stream, int k, const int32_t* lens_dev, int max_len, int t_begin, int t_end,
                                           const double* traj, int64_t ld, cck_conflict_run* out_dev, unsigned long long* key_out_dev) {
    if (!c) return CCK_ERR_INVALID;
    if (!c->has_pvc) return fail(c, CCK_ERR_INVALID, "cck_set_pvc has not been called");
    if (k != c->pvc.k) return fail(c, CCK_ERR_INVALID, "k differs from cck_set_pvc");
    if (max_len < 1 || ld < max_len) return fail(c, CCK_ERR_INVALID, "bad max_len/ld");
    if (t_begin < 0 || t_end > max_len || t_begin > t_end) return fail(c, CCK_ERR_INVALID, "bad step range");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = pick(c, stream);
    CK(c, cudaMemsetAsync(c->d_key, 0xff, sizeof(unsigned long long), s));
    if (c->pvc.n_pairs > 0 && t_end > t_begin) {
        const int warps_per_block = 4;
        int blocks = (t_end - t_begin + warps_per_block - 1) / warps_per_block;
        if (blocks > c->num_sms * 8) blocks = c->num_sms * 8;
        k_validate_steps<<<blocks, 128, 0, s>>>(c->pvc, traj, ld, lens_dev, t_begin, t_end, c->d_key);
        c->launches++;
    }
    k_validate_finalize<<<1, 256, 0, s>>>(c->pvc, traj, ld, lens_dev, max_len, c->d_key, out_dev);
    c->launches++;
    if (key_out_dev) CK(c, cudaMemcpyAsync(key_out_dev, c->d_key, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
    CK(c, cudaGetLastError());
    return CCK_OK;
}

extern "C" int cck_validate_plan_dev(cck_ctx* c, void* stream, int k, const int32_t* lens_dev, int max_len, const double* traj,
                                     int64_t ld, cck_conflict_run* out_dev) {
    return cck_validate_plan_range_dev(c, stream, k, lens_dev, max_len, 0, max_len, traj, ld, out_dev, nullptr);
}

// ---------------------------------------------------------------------------
// constraints
// ---------------------------------------------------------------------------
extern "C" int cck_set_constraints(cck_ctx* c, int n_entries, const int32_t* ent_step, const int32_t* pair_ab, const double* mu2,
                                   const double* cov4) {
    if (!c) return CCK_ERR_INVALID;
    if (!c->has_pvc) return fail(c, CCK_ERR_INVALID, "cck_set_pvc has not been called");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    if (n_entries <= 0) { cudaFree(c->d_cons); c->d_cons = nullptr; c->cons_cap = 0; c->cons_bytes = 0; return CCK_OK; }
    if (!ent_step || !pair_ab || !mu2 || !cov4) return fail(c, CCK_ERR_INVALID, "null constraint arrays");
    int max_step = 0;
    for (int e = 0; e < n_entries; ++e) {
        if (ent_step[e] < 0) return fail(c, CCK_ERR_INVALID, "negative constraint step");
        if (pair_ab[e] < 0 || pair_ab[e] >= c->pvc.k * c->pvc.k) return fail(c, CCK_ERR_INVALID, "pair index out of range");
        if (pair_ab[e] / c->pvc.k == pair_ab[e] % c->pvc.k) return fail(c, CCK_ERR_INVALID, "a robot cannot constrain itself");
        max_step = std::max(max_step, ent_step[e]);
    }
    std::vector<int> order(n_entries);
    for (int e = 0; e < n_entries; ++e) order[e] = e;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return ent_step[x] < ent_step[y]; });
    ConsHeader h{};
    h.n_entries = n_entries; h.max_step = max_step;
    h.off_first = sizeof(ConsHeader);
    h.off_entries = (h.off_first + (max_step + 2) * 4 + 15) & ~15;
    h.pvc_kind = c->pvc.kind; h.dim = c->pvc.dim; h.k = c->pvc.k; h.grid_n = c->pvc.grid_n;
    h.c_pair = c->pvc.c_pair; h.p_coll_agnts = c->pvc.p_coll_agnts; h.sc = c->pvc.sc;
    const size_t total = h.off_entries + (size_t)n_entries * sizeof(ConsEntry);
    std::vector<unsigned char> buf(total, 0);
    std::memcpy(buf.data(), &h, sizeof(h));
    int32_t* first = reinterpret_cast<int32_t*>(buf.data() + h.off_first);
    ConsEntry* ent = reinterpret_cast<ConsEntry*>(buf.data() + h.off_entries);
    int e = 0;
    for (int s = 0; s <= max_step + 1; ++s) {
        while (e < n_entries && ent_step[order[e]] < s) ++e;
        first[s] = e;
    }
    for (int i = 0; i < n_entries; ++i) {
        const int src = order[i];
        std::memcpy(ent[i].mu, mu2 + 2 * src, 16);
        std::memcpy(ent[i].cov, cov4 + 4 * src, 32);
        // pair_ab = constrained * k + constraining; the half-planes belong to the unordered pair (stored at [lo*k + hi])
        const int k = c->pvc.k, ra = pair_ab[src] / k, rb = pair_ab[src] % k;
        const size_t hp = (size_t)std::min(ra, rb) * k + std::max(ra, rb);
        std::memcpy(ent[i].A, c->pvc_A.data() + hp * 8, 64);
        std::memcpy(ent[i].B, c->pvc_B.data() + hp * 4, 32);
        ent[i].rad[0] = c->pvc_radii[ra]; ent[i].rad[1] = c->pvc_radii[rb];
    }
    if (total > c->cons_cap) {
        cudaFree(c->d_cons); c->d_cons = nullptr; c->cons_cap = 0;
        CK(c, cudaMalloc(&c->d_cons, total + 16));
        c->cons_cap = total;
    }
    CK(c, cudaMemcpy(c->d_cons, buf.data(), total, cudaMemcpyHostToDevice));
    c->cons_bytes = (int)((total + 15) & ~(size_t)15);
    return CCK_OK;
}

extern "C" int cck_satisfies_constraints_dev(cck_ctx* c, void* stream, int64_t n_paths, const double* path, int64_t ld, int32_t max_len,
                                             const int32_t* path_len, const int32_t* path_step0, uint8_t* ok) {
    if (!c) return CCK_ERR_INVALID;
    if (!c->has_pvc) return fail(c, CCK_ERR_INVALID, "cck_set_pvc has not been called");
    if (n_paths < 0 || ld < n_paths || max_len < 0) return fail(c, CCK_ERR_INVALID, "bad n/ld");
    if (n_paths == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = pick(c, stream);
    if (!c->d_cons) {   // no constraints: every path passes
        CK(c, cudaMemsetAsync(ok, 1, (size_t)n_paths, s));
        return CCK_OK;
    }
    k_constraints<<<grid_for(c, n_paths, 128, 8), 128, 0, s>>>(n_paths, c->pvc.dim, path, ld, path_len, path_step0, c->d_cons, ok);
    c->launches++;
    CK(c, cudaGetLastError());
    return CCK_OK;
}

// ---------------------------------------------------------------------------
// host-pointer entry points: H2D -> kernel -> D2H, chunked over two pipeline
// slots/streams so copies of chunk i+1 overlap the kernel of chunk i.
// ---------------------------------------------------------------------------
static int ensure_scratch(cck_ctx* c, int slot, size_t bytes) {
    if (bytes <= c->scratch_cap[slot]) return CCK_OK;
    CK(c, cudaStreamSynchronize(c->pipe[slot]));
    cudaFree(c->d_scratch[slot]); c->d_scratch[slot] = nullptr; c->scratch_cap[slot] = 0;
    const size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&c->d_scratch[slot], want);
    if (e != cudaSuccess) { c->err = std::string("cudaMalloc scratch: ") + cudaGetErrorString(e); return CCK_ERR_NOMEM; }
    c->scratch_cap[slot] = want;
    return CCK_OK;
}
static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// copy `planes` planes of `cnt` elements: host (pitch ld) <-> device (pitch cnt)
// SoA planes host <-> device.  Contiguous planes (ld == cnt: what the planner passes) go as ONE copy: a 2-D copy from
// pageable memory is staged row by row by the driver (36 planes of a unicycle batch cost ~1 ms instead of ~40 us).
static cudaError_t h2d_planes(void* dst, const void* src, size_t elem, int64_t cnt, int64_t ld_src, int planes, cudaStream_t s) {
    if (ld_src == cnt) return cudaMemcpyAsync(dst, src, (size_t)cnt * elem * planes, cudaMemcpyHostToDevice, s);
    return cudaMemcpy2DAsync(dst, (size_t)cnt * elem, src, (size_t)ld_src * elem, (size_t)cnt * elem, planes, cudaMemcpyHostToDevice, s);
}
static cudaError_t d2h_planes(void* dst, int64_t ld_dst, const void* src, size_t elem, int64_t cnt, int planes, cudaStream_t s) {
    if (ld_dst == cnt) return cudaMemcpyAsync(dst, src, (size_t)cnt * elem * planes, cudaMemcpyDeviceToHost, s);
    return cudaMemcpy2DAsync(dst, (size_t)ld_dst * elem, src, (size_t)cnt * elem, (size_t)cnt * elem, planes, cudaMemcpyDeviceToHost, s);
}


// CentralizedChiSquaredBoundarySVC's agent-agent decagon test for n pairs (host pointers; SoA: mu [2][n], cov [4][n])
extern "C" int cck_chisq_pairs_disjoint(cck_ctx* c, int64_t n, const double* mu_a, const double* cov_a, const double* rad_a, const double* mu_b,
                                        const double* cov_b, const double* rad_b, double sc, uint8_t* out) {
    if (!c) return CCK_ERR_INVALID;
    if (n < 0 || (n > 0 && (!mu_a || !cov_a || !rad_a || !mu_b || !cov_b || !rad_b || !out))) return fail(c, CCK_ERR_INVALID, "bad arguments");
    if (n == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    int rc;
    const size_t b1 = al256((size_t)n * 8), total = 14 * b1 + al256((size_t)n);
    if ((rc = ensure_scratch(c, 0, total))) return rc;
    unsigned char* p = (unsigned char*)c->d_scratch[0];
    cudaStream_t s = c->pipe[0];
    double* d_mua = (double*)p; double* d_cova = (double*)(p + 2 * b1); double* d_rada = (double*)(p + 6 * b1);
    double* d_mub = (double*)(p + 7 * b1); double* d_covb = (double*)(p + 9 * b1); double* d_radb = (double*)(p + 13 * b1);
    uint8_t* d_out = p + 14 * b1;
    // planes are packed at pitch n inside the al256-sized slots (2 resp. 4 planes each fit: al256(n*8) >= n*8)
    CK(c, cudaMemcpyAsync(d_mua, mu_a, (size_t)n * 16, cudaMemcpyHostToDevice, s)); CK(c, cudaMemcpyAsync(d_cova, cov_a, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(d_rada, rad_a, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(d_mub, mu_b, (size_t)n * 16, cudaMemcpyHostToDevice, s)); CK(c, cudaMemcpyAsync(d_covb, cov_b, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(d_radb, rad_b, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    SvcDev sp = c->svc;
    const double two_pi = 2.0 * M_PI, diff = two_pi / 10.0;      // point_circle(10), as in cck_set_obstacles
    double ang = 0;
    for (int i = 0; i < 10; ++i, ang -= diff) { sp.dec_cos[i] = std::cos(ang); sp.dec_sin[i] = std::sin(ang); }
    k_chisq_pairs<<<grid_for(c, n, 128, 8), 128, 0, s>>>(n, d_mua, d_cova, d_rada, d_mub, d_covb, d_radb, sc, sp, d_out);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, d_out, (size_t)n, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return CCK_OK;
}

extern "C" int cck_propagate(cck_ctx* c, int64_t n, const double* bel, int64_t ld, const double* ctrl, int64_t ldc, double duration,
                             double* out, int64_t ldo) {
    int rc = check_ready(c, false);
    if (rc) return rc;
    if (n < 0 || ld < n || ldc < n || ldo < n) return fail(c, CCK_ERR_INVALID, "bad n/ld");
    CK(c, cudaSetDevice(c->device));
    const int dim = c->model.dim, W = dim + 2 * dim * dim;
    int slot = 0;
    for (int64_t off = 0; off < n; off += kChunk, slot = (slot + 1) % CCK_SLOTS) {
        const int64_t cnt = std::min(kChunk, n - off);
        const size_t b_in = al256((size_t)cnt * W * 8), b_c = al256((size_t)cnt * dim * 8), b_out = b_in;
        if ((rc = ensure_scratch(c, slot, b_in + b_c + b_out))) return rc;
        unsigned char* base = (unsigned char*)c->d_scratch[slot];
        double* d_in = (double*)base; double* d_c = (double*)(base + b_in); double* d_out = (double*)(base + b_in + b_c);
        cudaStream_t s = c->pipe[slot];
        CK(c, h2d_planes(d_in, bel + off, 8, cnt, ld, W, s));
        CK(c, h2d_planes(d_c, ctrl + off, 8, cnt, ldc, dim, s));
        if ((rc = cck_propagate_dev(c, s, cnt, d_in, cnt, d_c, cnt, duration, d_out, cnt))) return rc;
        CK(c, d2h_planes(out + off, ldo, d_out, 8, cnt, W, s));
    }
    for (int i = 0; i < CCK_SLOTS; ++i) CK(c, cudaStreamSynchronize(c->pipe[i]));
    return CCK_OK;
}

extern "C" int cck_is_valid(cck_ctx* c, int64_t n, const double* bel, int64_t ld, uint8_t* valid) {
    int rc = check_ready(c, true);
    if (rc) return rc;
    if (n < 0 || ld < n) return fail(c, CCK_ERR_INVALID, "bad n/ld");
    CK(c, cudaSetDevice(c->device));
    const int dim = c->model.dim, W = dim + 2 * dim * dim;
    int slot = 0;
    for (int64_t off = 0; off < n; off += kChunk, slot = (slot + 1) % CCK_SLOTS) {
        const int64_t cnt = std::min(kChunk, n - off);
        const size_t b_in = al256((size_t)cnt * W * 8), b_v = al256((size_t)cnt);
        if ((rc = ensure_scratch(c, slot, b_in + b_v))) return rc;
        unsigned char* base = (unsigned char*)c->d_scratch[slot];
        cudaStream_t s = c->pipe[slot];
        CK(c, h2d_planes(base, bel + off, 8, cnt, ld, W, s));
        if ((rc = cck_is_valid_dev(c, s, cnt, (double*)base, cnt, base + b_in))) return rc;
        CK(c, cudaMemcpyAsync(valid + off, base + b_in, (size_t)cnt, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < CCK_SLOTS; ++i) CK(c, cudaStreamSynchronize(c->pipe[i]));
    return CCK_OK;
}

// Planner-sized batches (n <= 16384, no trajectory output): the ~14 small copies of the chunked path cost more
// than the launch itself, so everything is packed into ONE pinned staging block: one H2D, one launch, one D2H.
static int rollout_host_small(cck_ctx* c, bool expand, int64_t n, const double* bel, int64_t ld, const double* ctrl, int64_t ldc,
                              const int32_t* steps, const int32_t* parent_step, const double* parent_cost, const cck_goal_params* goal,
                              double* out, int64_t ldo, int32_t* valid_steps, double* cost, double* goal_dist, uint8_t* flags) {
    int rc;
    const int dim = c->model.dim, W = dim + 2 * dim * dim;
    const size_t b_in = al256((size_t)n * W * 8), b_c = al256((size_t)n * dim * 8), b_i = al256((size_t)n * 4), b_d = al256((size_t)n * 8), b_f = al256((size_t)n);
    const size_t in_bytes = b_in + b_c + b_i + (expand ? b_i + b_d : 0);
    const size_t out_bytes = b_in + b_i + (expand ? 2 * b_d + b_f : 0);
    const size_t total = in_bytes + out_bytes;
    if (total > c->h_stage_cap) {
        CK(c, cudaStreamSynchronize(c->pipe[0]));
        cudaFreeHost(c->h_stage); c->h_stage = nullptr; c->h_stage_cap = 0;
        if (cudaHostAlloc((void**)&c->h_stage, total * 2, cudaHostAllocDefault) != cudaSuccess) { c->err = "cudaHostAlloc staging"; return CCK_ERR_NOMEM; }
        c->h_stage_cap = total * 2;
    }
    if ((rc = ensure_scratch(c, 0, total))) return rc;
    unsigned char* h = c->h_stage;
    unsigned char* d = (unsigned char*)c->d_scratch[0];
    size_t o = 0;
    const size_t o_bel = o; o += b_in;
    const size_t o_ctrl = o; o += b_c;
    const size_t o_steps = o; o += b_i;
    size_t o_ps = 0, o_pc = 0;
    if (expand) { o_ps = o; o += b_i; o_pc = o; o += b_d; }
    const size_t o_out = o; o += b_in;
    const size_t o_valid = o; o += b_i;
    size_t o_cost = 0, o_gd = 0, o_fl = 0;
    if (expand) { o_cost = o; o += b_d; o_gd = o; o += b_d; o_fl = o; o += b_f; }
    for (int f = 0; f < W; ++f) std::memcpy(h + o_bel + (size_t)f * n * 8, bel + (size_t)f * ld, (size_t)n * 8);
    for (int f = 0; f < dim; ++f) std::memcpy(h + o_ctrl + (size_t)f * n * 8, ctrl + (size_t)f * ldc, (size_t)n * 8);
    std::memcpy(h + o_steps, steps, (size_t)n * 4);
    if (expand) { std::memcpy(h + o_ps, parent_step, (size_t)n * 4); std::memcpy(h + o_pc, parent_cost, (size_t)n * 8); }
    cudaStream_t s = c->pipe[0];
    CK(c, cudaMemcpyAsync(d, h, in_bytes, cudaMemcpyHostToDevice, s));
    if (!expand) rc = cck_propagate_while_valid_dev(c, s, n, (double*)(d + o_bel), n, (double*)(d + o_ctrl), n, (int32_t*)(d + o_steps), (double*)(d + o_out), n,
                                                    (int32_t*)(d + o_valid), nullptr, 0, 0);
    else rc = cck_expand_batch_dev(c, s, n, (double*)(d + o_bel), n, (double*)(d + o_ctrl), n, (int32_t*)(d + o_steps), (int32_t*)(d + o_ps), (double*)(d + o_pc),
                                   goal, (double*)(d + o_out), n, (int32_t*)(d + o_valid), (double*)(d + o_cost), (double*)(d + o_gd), d + o_fl);
    if (rc) return rc;
    CK(c, cudaMemcpyAsync(h + in_bytes, d + in_bytes, out_bytes, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    for (int f = 0; f < W; ++f) std::memcpy(out + (size_t)f * ldo, h + o_out + (size_t)f * n * 8, (size_t)n * 8);
    std::memcpy(valid_steps, h + o_valid, (size_t)n * 4);
    if (expand) { std::memcpy(cost, h + o_cost, (size_t)n * 8); std::memcpy(goal_dist, h + o_gd, (size_t)n * 8); std::memcpy(flags, h + o_fl, (size_t)n); }
    return CCK_OK;
}

static int rollout_host(cck_ctx* c, bool expand, int64_t n, const double* bel, int64_t ld, const double* ctrl, int64_t ldc,
                        const int32_t* steps, const int32_t* parent_step, const double* parent_cost, const cck_goal_params* goal,
                        double* out, int64_t ldo, int32_t* valid_steps, double* traj, int64_t traj_ld, int32_t traj_T, double* cost,
                        double* goal_dist, uint8_t* flags) {
    int rc = check_ready(c, true);
    if (rc) return rc;
    if (n < 0 || ld < n || ldc < n || ldo < n || (traj && (traj_ld < n || traj_T < 1))) return fail(c, CCK_ERR_INVALID, "bad n/ld");
    if (!bel || !ctrl || !steps || !out || !valid_steps) return fail(c, CCK_ERR_INVALID, "null argument");
    if (traj) {   // the trajectory buffer holds traj_T steps per belief: a longer rollout would write past it
        int32_t mx = 0;
        for (int64_t i = 0; i < n; ++i) mx = std::max(mx, steps[i]);
        if (mx > traj_T) return fail(c, CCK_ERR_INVALID, "max(steps) = " + std::to_string(mx) + " exceeds traj_T = " + std::to_string(traj_T));
    }
    CK(c, cudaSetDevice(c->device));
    if (n > 0 && n <= 16384 && !traj)
        return rollout_host_small(c, expand, n, bel, ld, ctrl, ldc, steps, parent_step, parent_cost, goal, out, ldo, valid_steps, cost, goal_dist, flags);
    const int dim = c->model.dim, W = dim + 2 * dim * dim;
    int64_t chunk = kChunk;
    if (traj) chunk = std::max<int64_t>(1024, kChunk / (traj_T > 0 ? traj_T : 1));
    int slot = 0;
    for (int64_t off = 0; off < n; off += chunk, slot = (slot + 1) % CCK_SLOTS) {
        const int64_t cnt = std::min(chunk, n - off);
        const size_t b_in = al256((size_t)cnt * W * 8), b_c = al256((size_t)cnt * dim * 8), b_i = al256((size_t)cnt * 4);
        const size_t b_d = al256((size_t)cnt * 8), b_f = al256((size_t)cnt);
        const size_t b_traj = traj ? al256((size_t)cnt * W * 8 * traj_T) : 0;
        size_t total = b_in + b_c + b_i + b_in + b_i + b_traj;
        if (expand) total += b_i + b_d + b_d + b_d + b_f;
        if ((rc = ensure_scratch(c, slot, total))) return rc;
        unsigned char* p = (unsigned char*)c->d_scratch[slot];
        double* d_in = (double*)p; p += b_in;
        double* d_c = (double*)p; p += b_c;
        int32_t* d_steps = (int32_t*)p; p += b_i;
        double* d_out = (double*)p; p += b_in;
        int32_t* d_valid = (int32_t*)p; p += b_i;
        double* d_traj = traj ? (double*)p : nullptr; p += b_traj;
        cudaStream_t s = c->pipe[slot];
        CK(c, h2d_planes(d_in, bel + off, 8, cnt, ld, W, s));
        CK(c, h2d_planes(d_c, ctrl + off, 8, cnt, ldc, dim, s));
        CK(c, cudaMemcpyAsync(d_steps, steps + off, (size_t)cnt * 4, cudaMemcpyHostToDevice, s));
        if (!expand) {
            if ((rc = cck_propagate_while_valid_dev(c, s, cnt, d_in, cnt, d_c, cnt, d_steps, d_out, cnt, d_valid, d_traj, cnt, traj_T))) return rc;
        } else {
            int32_t* d_ps = (int32_t*)p; p += b_i;
            double* d_pc = (double*)p; p += b_d;
            double* d_cost = (double*)p; p += b_d;
            double* d_gd = (double*)p; p += b_d;
            uint8_t* d_fl = (uint8_t*)p; p += b_f;
            CK(c, cudaMemcpyAsync(d_ps, parent_step + off, (size_t)cnt * 4, cudaMemcpyHostToDevice, s));
            CK(c, cudaMemcpyAsync(d_pc, parent_cost + off, (size_t)cnt * 8, cudaMemcpyHostToDevice, s));
            if ((rc = cck_expand_batch_dev(c, s, cnt, d_in, cnt, d_c, cnt, d_steps, d_ps, d_pc, goal, d_out, cnt, d_valid, d_cost, d_gd, d_fl))) return rc;
            CK(c, cudaMemcpyAsync(cost + off, d_cost, (size_t)cnt * 8, cudaMemcpyDeviceToHost, s));
            CK(c, cudaMemcpyAsync(goal_dist + off, d_gd, (size_t)cnt * 8, cudaMemcpyDeviceToHost, s));
            CK(c, cudaMemcpyAsync(flags + off, d_fl, (size_t)cnt, cudaMemcpyDeviceToHost, s));
        }
        CK(c, d2h_planes(out + off, ldo, d_out, 8, cnt, W, s));
        CK(c, cudaMemcpyAsync(valid_steps + off, d_valid, (size_t)cnt * 4, cudaMemcpyDeviceToHost, s));
        if (traj) CK(c, d2h_planes(traj + off, traj_ld, d_traj, 8, cnt, W * traj_T, s));
    }
    for (int i = 0; i < CCK_SLOTS; ++i) CK(c, cudaStreamSynchronize(c->pipe[i]));
    return CCK_OK;
}

extern "C" int cck_propagate_while_valid(cck_ctx* c, int64_t n, const double* bel, int64_t ld, const double* ctrl, int64_t ldc,
                                         const int32_t* steps, double* out, int64_t ldo, int32_t* valid_steps, double* traj,
                                         int64_t traj_ld, int32_t traj_T) {
    return rollout_host(c, false, n, bel, ld, ctrl, ldc, steps, nullptr, nullptr, nullptr, out, ldo, valid_steps, traj, traj_ld, traj_T,
                        nullptr, nullptr, nullptr);
}

extern "C" int cck_expand_batch(cck_ctx* c, int64_t n, const double* bel, int64_t ld, const double* ctrl, int64_t ldc,
                                const int32_t* steps, const int32_t* parent_step, const double* parent_cost, const cck_goal_params* goal,
                                double* out, int64_t ldo, int32_t* valid_steps, double* cost, double* goal_dist, uint8_t* flags) {
    if (!c) return CCK_ERR_INVALID;
    if (!goal || !parent_step || !parent_cost || !cost || !goal_dist || !flags) return fail(c, CCK_ERR_INVALID, "null argument");
    return rollout_host(c, true, n, bel, ld, ctrl, ldc, steps, parent_step, parent_cost, goal, out, ldo, valid_steps, nullptr, 0, 0, cost,
                        goal_dist, flags);
}

// t_end < 0: the whole plan.  key_out (optional): (first_step << 32 | rank of the pair in map order), ~0 = no conflict in the range.
extern "C" int cck_validate_plan_range(cck_ctx* c, int k, const int32_t* lens, const double* traj, int64_t ld, int t_begin, int t_end,
                                       cck_conflict_run* out, unsigned long long* key_out) {
    if (!c) return CCK_ERR_INVALID;
    if (!c->has_pvc) return fail(c, CCK_ERR_INVALID, "cck_set_pvc has not been called");
    if (!lens || !traj || !out || k != c->pvc.k) return fail(c, CCK_ERR_INVALID, "bad arguments");
    CK(c, cudaSetDevice(c->device));
    int max_len = 0;
    for (int r = 0; r < k; ++r) { if (lens[r] < 1) return fail(c, CCK_ERR_INVALID, "empty trajectory"); max_len = std::max(max_len, lens[r]); }
    if (ld < max_len) return fail(c, CCK_ERR_INVALID, "ld < max trajectory length");
    if (t_end < 0) t_end = max_len;
    if (t_end > max_len) t_end = max_len;
    if (t_begin > t_end) t_begin = t_end;
    const int W = c->pvc.dim + 2 * c->pvc.dim * c->pvc.dim;
    const size_t b_traj = al256((size_t)k * W * max_len * 8), b_len = al256((size_t)k * 4);
    int rc;
    if ((rc = ensure_scratch(c, 0, b_traj + b_len + 256))) return rc;
    unsigned char* p = (unsigned char*)c->d_scratch[0];
    cudaStream_t s = c->pipe[0];
    CK(c, h2d_planes(p, traj, 8, max_len, ld, k * W, s));
    CK(c, cudaMemcpyAsync(p + b_traj, lens, (size_t)k * 4, cudaMemcpyHostToDevice, s));
    unsigned long long* d_keyout = (unsigned long long*)(p + b_traj + b_len);
    if ((rc = cck_validate_plan_range_dev(c, s, k, (int32_t*)(p + b_traj), max_len, t_begin, t_end, (double*)p, max_len, c->d_run,
                                          key_out ? d_keyout : nullptr))) return rc;
    CK(c, cudaMemcpyAsync(out, c->d_run, sizeof(cck_conflict_run), cudaMemcpyDeviceToHost, s));
    if (key_out) CK(c, cudaMemcpyAsync(key_out, d_keyout, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return CCK_OK;
}

extern "C" int cck_validate_plan(cck_ctx* c, int k, const int32_t* lens, const double* traj, int64_t ld, cck_conflict_run* out) {
    return cck_validate_plan_range(c, k, lens, traj, ld, 0, -1, out, nullptr);
}

extern "C" int cck_satisfies_constraints(cck_ctx* c, int64_t n_paths, const double* path, int64_t ld, int32_t max_len,
                                         const int32_t* path_len, const int32_t* path_step0, uint8_t* ok) {
    if (!c) return CCK_ERR_INVALID;
    if (!c->has_pvc) return fail(c, CCK_ERR_INVALID, "cck_set_pvc has not been called");
    if (!path || !path_len || !path_step0 || !ok || n_paths < 0 || ld < n_paths || max_len < 1) return fail(c, CCK_ERR_INVALID, "bad arguments");
    if (n_paths == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    const int W = c->pvc.dim + 2 * c->pvc.dim * c->pvc.dim;
    const size_t b_path = al256((size_t)n_paths * W * max_len * 8), b_i = al256((size_t)n_paths * 4), b_ok = al256((size_t)n_paths);
    int rc;
    if ((rc = ensure_scratch(c, 0, b_path + 2 * b_i + b_ok))) return rc;
    unsigned char* p = (unsigned char*)c->d_scratch[0];
    cudaStream_t s = c->pipe[0];
    CK(c, h2d_planes(p, path, 8, n_paths, ld, W * max_len, s));
    CK(c, cudaMemcpyAsync(p + b_path, path_len, (size_t)n_paths * 4, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(p + b_path + b_i, path_step0, (size_t)n_paths * 4, cudaMemcpyHostToDevice, s));
    if ((rc = cck_satisfies_constraints_dev(c, s, n_paths, (double*)p, n_paths, max_len, (int32_t*)(p + b_path), (int32_t*)(p + b_path + b_i),
                                            p + b_path + 2 * b_i))) return rc;
    CK(c, cudaMemcpyAsync(ok, p + b_path + 2 * b_i, (size_t)n_paths, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return CCK_OK;
}

// ---------------------------------------------------------------------------
// device-resident SST tree / witness set with batched belief-space
// nearest-neighbour queries (RealVectorBeliefSpace.cpp:16-62, ConstraintRespectingBSST.cpp:119-178)
// ---------------------------------------------------------------------------
struct cck_tree {
    cck_ctx* c = nullptr;
    int dim = 2;
    int64_t cap = 0, n = 0;
    double* planes = nullptr;
    uint8_t* active = nullptr;
    void* scratch = nullptr;
    size_t scratch_cap = 0;
};

static int tree_planes_of(int dim) { return dim + 2 * dim * dim + 1; }

static int tree_scratch(cck_tree* t, size_t bytes) {
    if (bytes <= t->scratch_cap) return CCK_OK;
    cck_ctx* c = t->c;
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(t->scratch); t->scratch = nullptr; t->scratch_cap = 0;
    cudaError_t e = cudaMalloc(&t->scratch, bytes + bytes / 2);
    if (e != cudaSuccess) { c->err = std::string("cudaMalloc tree scratch: ") + cudaGetErrorString(e); return CCK_ERR_NOMEM; }
    t->scratch_cap = bytes + bytes / 2;
    return CCK_OK;
}

static int tree_reserve(cck_tree* t, int64_t want) {
    if (want <= t->cap) return CCK_OK;
    cck_ctx* c = t->c;
    int64_t ncap = std::max<int64_t>(4096, t->cap * 2);
    while (ncap < want) ncap *= 2;
    const int P = tree_planes_of(t->dim);
    double* np = nullptr; uint8_t* na = nullptr;
    if (cudaMalloc(&np, (size_t)P * ncap * 8) != cudaSuccess || cudaMalloc(&na, (size_t)ncap) != cudaSuccess) {
        cudaFree(np); c->err = "cudaMalloc tree"; return CCK_ERR_NOMEM;
    }
    if (t->n > 0) {
        CK(c, cudaMemcpy2DAsync(np, (size_t)ncap * 8, t->planes, (size_t)t->cap * 8, (size_t)t->n * 8, P, cudaMemcpyDeviceToDevice, c->stream));
        CK(c, cudaMemcpyAsync(na, t->active, (size_t)t->n, cudaMemcpyDeviceToDevice, c->stream));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    cudaFree(t->planes); cudaFree(t->active);
    t->planes = np; t->active = na; t->cap = ncap;
    return CCK_OK;
}

extern "C" int cck_tree_create(cck_ctx* c, int dim, cck_tree** out) {
    if (!c || !out) return CCK_ERR_INVALID;
    *out = nullptr;
    if (dim != 2 && dim != 4) return fail(c, CCK_ERR_INVALID, "tree dim must be 2 or 4");
    cck_tree* t = new cck_tree();
    t->c = c; t->dim = dim;
    *out = t;
    return CCK_OK;
}
extern "C" int cck_tree_destroy(cck_tree* t) {
    if (!t) return CCK_OK;
    cudaSetDevice(t->c->device);
    cudaStreamSynchronize(t->c->stream);
    cudaFree(t->planes); cudaFree(t->active); cudaFree(t->scratch);
    delete t;
    return CCK_OK;
}
extern "C" int cck_tree_clear(cck_tree* t) { if (!t) return CCK_ERR_INVALID; t->n = 0; return CCK_OK; }
extern "C" int64_t cck_tree_size(const cck_tree* t) { return t ? t->n : 0; }

extern "C" int cck_tree_append(cck_tree* t, int64_t n, const double* bel, int64_t ld, const double* cost, int64_t* first) {
    if (!t) return CCK_ERR_INVALID;
    cck_ctx* c = t->c;
    if (n < 0 || ld < n || (n > 0 && (!bel || !cost))) return fail(c, CCK_ERR_INVALID, "bad tree append arguments");
    if (first) *first = t->n;
    if (n == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    int rc;
    if ((rc = tree_reserve(t, t->n + n))) return rc;
    const int W = t->dim + 2 * t->dim * t->dim;
    const size_t b_bel = al256((size_t)n * W * 8);
    if ((rc = tree_scratch(t, b_bel + al256((size_t)n * 8)))) return rc;
    double* d_bel = (double*)t->scratch; double* d_cost = (double*)((unsigned char*)t->scratch + b_bel);
    CK(c, h2d_planes(d_bel, bel, 8, n, ld, W, c->stream));
    CK(c, cudaMemcpyAsync(d_cost, cost, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    TreeDev td{t->planes, t->active, t->cap};
    const int g = grid_for(c, n, 128, 8);
    if (t->dim == 2) k_tree_append<2><<<g, 128, 0, c->stream>>>(td, t->n, n, d_bel, n, d_cost);
    else k_tree_append<4><<<g, 128, 0, c->stream>>>(td, t->n, n, d_bel, n, d_cost);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(c->stream));    // the host buffers may be reused by the caller
    t->n += n;
    return CCK_OK;
}

extern "C" int cck_tree_update(cck_tree* t, int64_t n, const int32_t* idx, const double* cost, const uint8_t* active) {
    if (!t) return CCK_ERR_INVALID;
    cck_ctx* c = t->c;
    if (n < 0 || (n > 0 && !idx)) return fail(c, CCK_ERR_INVALID, "bad tree update arguments");
    if (n == 0 || (!cost && !active)) return CCK_OK;
    for (int64_t j = 0; j < n; ++j) if (idx[j] < 0 || idx[j] >= t->n) return fail(c, CCK_ERR_INVALID, "tree update index out of range");
    CK(c, cudaSetDevice(c->device));
    int rc;
    const size_t b_i = al256((size_t)n * 4), b_c = al256((size_t)n * 8), b_a = al256((size_t)n);
    if ((rc = tree_scratch(t, b_i + b_c + b_a))) return rc;
    unsigned char* p = (unsigned char*)t->scratch;
    int32_t* d_idx = (int32_t*)p; double* d_cost = (double*)(p + b_i); uint8_t* d_act = p + b_i + b_c;
    CK(c, cudaMemcpyAsync(d_idx, idx, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    if (cost) CK(c, cudaMemcpyAsync(d_cost, cost, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    if (active) CK(c, cudaMemcpyAsync(d_act, active, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    TreeDev td{t->planes, t->active, t->cap};
    k_tree_update<<<grid_for(c, n, 128, 8), 128, 0, c->stream>>>(td, t->dim + 2 * t->dim * t->dim, n, d_idx, cost ? d_cost : nullptr, active ? d_act : nullptr);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaStreamSynchronize(c->stream));
    return CCK_OK;
}

static int tree_query(cck_tree* t, int mode, int64_t K, const double* q, int64_t ldq, double radius, int32_t* out_idx, double* out_dist) {
    if (!t) return CCK_ERR_INVALID;
    cck_ctx* c = t->c;
    if (K < 0 || ldq < K || (K > 0 && (!q || !out_idx))) return fail(c, CCK_ERR_INVALID, "bad tree query arguments");
    if (K == 0) return CCK_OK;
    if (K > INT_MAX) return fail(c, CCK_ERR_INVALID, "too many queries");
    CK(c, cudaSetDevice(c->device));
    int rc;
    const int W = t->dim + 2 * t->dim * t->dim;
    const size_t b_q = al256((size_t)K * W * 8), b_i = al256((size_t)K * 4), b_d = al256((size_t)K * 8);
    if ((rc = tree_scratch(t, b_q + b_i + b_d))) return rc;
    unsigned char* p = (unsigned char*)t->scratch;
    double* d_q = (double*)p; int32_t* d_idx = (int32_t*)(p + b_q); double* d_dist = (double*)(p + b_q + b_i);
    CK(c, h2d_planes(d_q, q, 8, K, ldq, W, c->stream));
    TreeDev td{t->planes, t->active, t->cap};
    if (t->dim == 2) {
        if (mode == 0) k_tree_query<2, 0><<<(int)K, 128, 0, c->stream>>>(td, t->n, d_q, K, radius, d_idx, d_dist);
        else k_tree_query<2, 1><<<(int)K, 128, 0, c->stream>>>(td, t->n, d_q, K, radius, d_idx, d_dist);
    } else {
        if (mode == 0) k_tree_query<4, 0><<<(int)K, 128, 0, c->stream>>>(td, t->n, d_q, K, radius, d_idx, d_dist);
        else k_tree_query<4, 1><<<(int)K, 128, 0, c->stream>>>(td, t->n, d_q, K, radius, d_idx, d_dist);
    }
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out_idx, d_idx, (size_t)K * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_dist) CK(c, cudaMemcpyAsync(out_dist, d_dist, (size_t)K * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return CCK_OK;
}

extern "C" int cck_tree_select(cck_tree* t, int64_t K, const double* samples, int64_t lds, double selection_radius, int32_t* out_idx,
                               double* out_dist) {
    return tree_query(t, 0, K, samples, lds, selection_radius, out_idx, out_dist);
}
extern "C" int cck_tree_nearest(cck_tree* t, int64_t K, const double* queries, int64_t ldq, int32_t* out_idx, double* out_dist) {
    return tree_query(t, 1, K, queries, ldq, 0.0, out_idx, out_dist);
}

extern "C" int cck_tree_witness_batch(cck_tree* t, int64_t n, const double* q, int64_t ldq, double pruning_radius, const uint8_t* eligible,
                                      int32_t* out_widx, double* out_wdist, int32_t* out_near, double* out_near_dist, uint8_t* out_fresh) {
    if (!t) return CCK_ERR_INVALID;
    cck_ctx* c = t->c;
    if (n < 0 || ldq < n || n > CCK_WIT_MAX || !(pruning_radius >= 0) ||
        (n > 0 && (!q || !out_widx || !out_wdist || !out_near || !out_near_dist || !out_fresh)))
        return fail(c, CCK_ERR_INVALID, "bad witness-batch arguments (n <= 1024)");
    if (n == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    int rc;
    const int W = t->dim + 2 * t->dim * t->dim, N = (int)n, nw = CCK_WIT_MAX / 32;
    // scratch: inputs and intermediates, then ONE result block [widx][near][wdist][near_dist][fresh] for a single D2H
    const size_t b_q = al256((size_t)n * W * 8), b_el = al256((size_t)n), b_flag = al256((size_t)n), b_lim = al256((size_t)n * 8),
                 b_D = al256((size_t)n * n * 8), b_adj = al256((size_t)n * nw * 4), b_fw = al256((size_t)nw * 4), b_cw = al256((size_t)nw * 4);
    const size_t r_i = al256((size_t)n * 4), r_d = al256((size_t)n * 8), r_f = al256((size_t)n), r_all = 2 * r_i + 2 * r_d + r_f;
    if ((rc = tree_scratch(t, b_q + b_el + 2 * b_flag + b_lim + b_D + b_adj + b_fw + b_cw + r_all))) return rc;
    if (c->h_stage_cap < r_all) {
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr; c->h_stage_cap = 0;
        if (cudaHostAlloc((void**)&c->h_stage, r_all * 2, cudaHostAllocDefault) != cudaSuccess) return fail(c, CCK_ERR_NOMEM, "cudaHostAlloc staging");
        c->h_stage_cap = r_all * 2;
    }
    unsigned char* p = (unsigned char*)t->scratch;
    double* d_q = (double*)p; p += b_q;
    uint8_t* d_el = (uint8_t*)p; p += b_el;
    uint8_t* d_iscol = (uint8_t*)p; p += b_flag;
    uint8_t* d_cand = (uint8_t*)p; p += b_flag;
    double* d_lim = (double*)p; p += b_lim;
    double* d_D = (double*)p; p += b_D;
    uint32_t* d_adj = (uint32_t*)p; p += b_adj;
    uint32_t* d_fw = (uint32_t*)p; p += b_fw;
    uint32_t* d_cw = (uint32_t*)p; p += b_cw;
    unsigned char* d_r = p;
    int32_t* d_widx = (int32_t*)d_r; int32_t* d_near = (int32_t*)(d_r + r_i);
    double* d_wdist = (double*)(d_r + 2 * r_i); double* d_ndist = (double*)(d_r + 2 * r_i + r_d);
    uint8_t* d_fresh = d_r + 2 * r_i + 2 * r_d;
    cudaStream_t s = c->stream;
    CK(c, h2d_planes(d_q, q, 8, n, ldq, W, s));
    if (eligible) CK(c, cudaMemcpyAsync(d_el, eligible, (size_t)n, cudaMemcpyHostToDevice, s));
    TreeDev td{t->planes, t->active, t->cap};
    const double r_up = std::nextafter(pruning_radius, HUGE_VAL);
    const size_t gsm = (size_t)n * nw * 4;
    if (t->dim == 2) {
        k_tree_query<2, 1><<<N, 128, 0, s>>>(td, t->n, d_q, n, 0.0, d_widx, d_wdist);
        k_wit_prep<<<CCK_WIT_MAX / 128, 128, 0, s>>>(N, d_widx, d_wdist, pruning_radius, r_up, eligible ? d_el : nullptr, d_iscol, d_cand, d_cw, d_lim);
        k_wit_rows<2><<<N, 128, 0, s>>>(N, d_q, n, d_cand, pruning_radius, d_D, d_adj);
    } else {
        k_tree_query<4, 1><<<N, 128, 0, s>>>(td, t->n, d_q, n, 0.0, d_widx, d_wdist);
        k_wit_prep<<<CCK_WIT_MAX / 128, 128, 0, s>>>(N, d_widx, d_wdist, pruning_radius, r_up, eligible ? d_el : nullptr, d_iscol, d_cand, d_cw, d_lim);
        k_wit_rows<4><<<N, 128, 0, s>>>(N, d_q, n, d_cand, pruning_radius, d_D, d_adj);
    }
    if ((rc = set_smem(c, k_wit_greedy, (int)gsm))) return rc;
    k_wit_greedy<<<1, 1024, gsm, s>>>(N, d_cw, d_adj, d_fw, d_fresh);
    k_wit_rowmin<<<(N + 3) / 4, 128, 0, s>>>(N, d_D, d_fw, d_lim, d_near, d_ndist);
    c->launches += 5;
    CK(c, cudaGetLastError());
    unsigned char* h = c->h_stage;
    CK(c, cudaMemcpyAsync(h, d_r, r_all, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    std::memcpy(out_widx, h, (size_t)n * 4); std::memcpy(out_near, h + r_i, (size_t)n * 4);
    std::memcpy(out_wdist, h + 2 * r_i, (size_t)n * 8); std::memcpy(out_near_dist, h + 2 * r_i + r_d, (size_t)n * 8);
    std::memcpy(out_fresh, h + 2 * r_i + 2 * r_d, (size_t)n);
    return CCK_OK;
}

extern "C" int cck_belief_distance_lower(cck_ctx* c, int dim, int64_t n, const double* bel, int64_t ld, double* out) {
    if (!c) return CCK_ERR_INVALID;
    if ((dim != 2 && dim != 4) || n < 0 || ld < n || n > 8192 || (n > 0 && (!bel || !out))) return fail(c, CCK_ERR_INVALID, "bad distance-matrix arguments (n <= 8192)");
    if (n == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    int rc;
    const int W = dim + 2 * dim * dim;
    const size_t b_in = al256((size_t)n * W * 8), b_o = al256((size_t)n * n * 8);
    if ((rc = ensure_scratch(c, 0, b_in + b_o))) return rc;
    unsigned char* p = (unsigned char*)c->d_scratch[0];
    double* d_b = (double*)p; double* d_o = (double*)(p + b_in);
    cudaStream_t s = c->pipe[0];
    CK(c, h2d_planes(d_b, bel, 8, n, ld, W, s));
    if (dim == 2) k_distance_lower<2><<<(int)n, 128, 0, s>>>(n, d_b, n, d_o);
    else k_distance_lower<4><<<(int)n, 128, 0, s>>>(n, d_b, n, d_o);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, d_o, (size_t)n * n * 8, cudaMemcpyDeviceToHost, s));   // entries with i >= q are unspecified
    CK(c, cudaStreamSynchronize(s));
    return CCK_OK;
}

extern "C" int cck_belief_distance_cols(cck_ctx* c, int dim, int64_t n, const double* bel, int64_t ld, int32_t nf, const int32_t* cols, double* out) {
    if (!c) return CCK_ERR_INVALID;
    if ((dim != 2 && dim != 4) || n < 0 || ld < n || nf < 0 || n > 65536 || (n > 0 && nf > 0 && (!bel || !cols || !out)))
        return fail(c, CCK_ERR_INVALID, "bad distance-columns arguments (n <= 65536)");
    if (n == 0 || nf == 0) return CCK_OK;
    for (int a = 0; a < nf; ++a) if (cols[a] < 0 || cols[a] >= n) return fail(c, CCK_ERR_INVALID, "distance column out of range");
    CK(c, cudaSetDevice(c->device));
    int rc;
    const int W = dim + 2 * dim * dim;
    const size_t b_in = al256((size_t)n * W * 8), b_c = al256((size_t)nf * 4), b_o = al256((size_t)n * nf * 8);
    if ((rc = ensure_scratch(c, 0, b_in + b_c + b_o))) return rc;
    unsigned char* p = (unsigned char*)c->d_scratch[0];
    double* d_b = (double*)p; int32_t* d_c = (int32_t*)(p + b_in); double* d_o = (double*)(p + b_in + b_c);
    cudaStream_t s = c->pipe[0];
    CK(c, h2d_planes(d_b, bel, 8, n, ld, W, s));
    CK(c, cudaMemcpyAsync(d_c, cols, (size_t)nf * 4, cudaMemcpyHostToDevice, s));
    if (dim == 2) k_distance_cols<2><<<(int)n, 128, 0, s>>>(n, d_b, n, nf, d_c, d_o);
    else k_distance_cols<4><<<(int)n, 128, 0, s>>>(n, d_b, n, nf, d_c, d_o);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, d_o, (size_t)n * nf * 8, cudaMemcpyDeviceToHost, s));   // entries with cols[a] >= q are unspecified
    CK(c, cudaStreamSynchronize(s));
    return CCK_OK;
}

extern "C" int cck_belief_distance_cols_near(cck_ctx* c, int dim, int64_t n, const double* bel, int64_t ld, int32_t nf, const int32_t* cols,
                                             const double* limit, int32_t cap, int32_t* out_idx, double* out_dist, int32_t* out_count) {
    if (!c) return CCK_ERR_INVALID;
    if ((dim != 2 && dim != 4) || n < 0 || ld < n || nf < 0 || n > 65536 || cap < 1 || cap > 65536 ||
        (n > 0 && (!out_count || (nf > 0 && (!bel || !cols || !limit || !out_idx || !out_dist)))))
        return fail(c, CCK_ERR_INVALID, "bad near-distance-columns arguments (n <= 65536, 1 <= cap <= 65536)");
    if (n == 0) return CCK_OK;
    if (nf == 0) { std::memset(out_count, 0, (size_t)n * 4); return CCK_OK; }
    for (int a = 0; a < nf; ++a)
        if (cols[a] < 0 || cols[a] >= n || (a > 0 && cols[a] <= cols[a - 1])) return fail(c, CCK_ERR_INVALID, "distance columns must be ascending and in range");
    CK(c, cudaSetDevice(c->device));
    int rc;
    const int W = dim + 2 * dim * dim;
    // device block: [bel planes][cols][limit] | result block (one D2H): [count n x i32][idx n x cap x i32][dist n x cap x f64]
    const size_t b_in = al256((size_t)n * W * 8), b_c = al256((size_t)nf * 4), b_l = al256((size_t)n * 8);
    const size_t r_cnt = al256((size_t)n * 4), r_idx = al256((size_t)n * cap * 4), r_dst = al256((size_t)n * cap * 8), r_all = r_cnt + r_idx + r_dst;
    if ((rc = ensure_scratch(c, 0, b_in + b_c + b_l + r_all))) return rc;
    if (c->h_stage_cap < r_all) {      // pinned landing zone for the result block (pageable D2H is several times slower)
        if (c->h_stage) cudaFreeHost(c->h_stage);
        c->h_stage = nullptr; c->h_stage_cap = 0;
        if (cudaHostAlloc((void**)&c->h_stage, r_all * 2, cudaHostAllocDefault) != cudaSuccess) return fail(c, CCK_ERR_NOMEM, "cudaHostAlloc staging");
        c->h_stage_cap = r_all * 2;
    }
    unsigned char* p = (unsigned char*)c->d_scratch[0];
    double* d_b = (double*)p; int32_t* d_c = (int32_t*)(p + b_in); double* d_l = (double*)(p + b_in + b_c);
    unsigned char* d_r = p + b_in + b_c + b_l;
    cudaStream_t s = c->pipe[0];
    CK(c, h2d_planes(d_b, bel, 8, n, ld, W, s));
    CK(c, cudaMemcpyAsync(d_c, cols, (size_t)nf * 4, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(d_l, limit, (size_t)n * 8, cudaMemcpyHostToDevice, s));
    const int g = (int)n;      // one CTA per row
    if (dim == 2) k_distance_cols_near<2><<<g, 128, 0, s>>>(n, d_b, n, nf, d_c, d_l, cap, (int32_t*)(d_r + r_cnt), (double*)(d_r + r_cnt + r_idx), (int32_t*)d_r);
    else k_distance_cols_near<4><<<g, 128, 0, s>>>(n, d_b, n, nf, d_c, d_l, cap, (int32_t*)(d_r + r_cnt), (double*)(d_r + r_cnt + r_idx), (int32_t*)d_r);
    c->launches++;
    CK(c, cudaGetLastError());
    unsigned char* h = (unsigned char*)c->h_stage;
    CK(c, cudaMemcpyAsync(h, d_r, r_all, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    std::memcpy(out_count, h, (size_t)n * 4);
    // only the filled prefix of every row is copied out (rows are mostly empty)
    const int32_t* hi = (const int32_t*)(h + r_cnt); const double* hd = (const double*)(h + r_cnt + r_idx);
    for (int64_t q = 0; q < n; ++q) {
        const int m = std::min(out_count[q], cap);
        if (m > 0) { std::memcpy(out_idx + q * cap, hi + q * cap, (size_t)m * 4); std::memcpy(out_dist + q * cap, hd + q * cap, (size_t)m * 8); }
    }
    return CCK_OK;
}

extern "C" int cck_belief_distance(cck_ctx* c, int dim, int64_t n, const double* a, int64_t lda, const double* b, int64_t ldb, double* out) {
    if (!c) return CCK_ERR_INVALID;
    if ((dim != 2 && dim != 4) || n < 0 || lda < n || ldb < n || (n > 0 && (!a || !b || !out))) return fail(c, CCK_ERR_INVALID, "bad distance arguments");
    if (n == 0) return CCK_OK;
    CK(c, cudaSetDevice(c->device));
    int rc;
    const int W = dim + 2 * dim * dim;
    const size_t b_in = al256((size_t)n * W * 8), b_o = al256((size_t)n * 8);
    if ((rc = ensure_scratch(c, 0, 2 * b_in + b_o))) return rc;
    unsigned char* p = (unsigned char*)c->d_scratch[0];
    double* d_a = (double*)p; double* d_b = (double*)(p + b_in); double* d_o = (double*)(p + 2 * b_in);
    cudaStream_t s = c->pipe[0];
    CK(c, h2d_planes(d_a, a, 8, n, lda, W, s));
    CK(c, h2d_planes(d_b, b, 8, n, ldb, W, s));
    if (dim == 2) k_belief_distance<2><<<grid_for(c, n, 128, 8), 128, 0, s>>>(n, d_a, n, d_b, n, d_o);
    else k_belief_distance<4><<<grid_for(c, n, 128, 8), 128, 0, s>>>(n, d_a, n, d_b, n, d_o);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, d_o, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return CCK_OK;
}

extern "C" int cck_rollout_summary_dev(cck_ctx* c, void* stream, int64_t n, const int32_t* steps, const int32_t* valid_steps,
                                       uint64_t* out_dev) {
    if (!c || !steps || !valid_steps || !out_dev || n < 0) return CCK_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = pick(c, stream);
    CK(c, cudaMemsetAsync(out_dev, 0, 3 * sizeof(uint64_t), s));
    if (n > 0) {
        k_summary<<<grid_for(c, n, 256, 4), 256, 0, s>>>(n, steps, valid_steps, reinterpret_cast<unsigned long long*>(out_dev));
        c->launches++;
        CK(c, cudaGetLastError());
    }
    return CCK_OK;
}

extern "C" int cck_select_winner_dev(cck_ctx* c, void* stream, int64_t n, const double* cost, const uint8_t* flags, int required_mask,
                                     int64_t index_offset, double* out_dev) {
    if (!c || !cost || !flags || !out_dev || n < 0) return CCK_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    k_winner<<<1, 1024, 0, pick(c, stream)>>>(n, cost, flags, required_mask, index_offset, out_dev);
    c->launches++;
    CK(c, cudaGetLastError());
    return CCK_OK;
}

// ---------------------------------------------------------------------------
// FP64 peak microbenchmark
// ---------------------------------------------------------------------------
extern "C" int cck_measure_fp64_peak(cck_ctx* c, double* tflops) {
    if (!c || !tflops) return CCK_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    double* d = nullptr;
    CK(c, cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    CK(c, cudaEventCreate(&e0)); CK(c, cudaEventCreate(&e1));
    const int iters = 1 << 15, blocks = c->num_sms * 8, threads = 256;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(c, cudaEventRecord(e0, c->stream));
        k_fp64_peak<<<blocks, threads, 0, c->stream>>>(d, iters, 1.0 + rep);
        CK(c, cudaEventRecord(e1, c->stream));
        CK(c, cudaEventSynchronize(e1));
        c->launches++;
        float ms = 0;
        CK(c, cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best;
    return CCK_OK;
}

// out[0..2]: FP64-pipe instructions per second (x 1e12, per lane) of alternating DMUL / DADD, DADD only, DMUL only;
// out[3..5]: the FP64 rate of the alternating stream with 1, 2, 3 independent fp32 FMULs issued per FP64 instruction
extern "C" int cck_measure_fp64_rates(cck_ctx* c, double* out) {
    if (!c || !out) return CCK_ERR_INVALID;
    CK(c, cudaSetDevice(c->device));
    double* d = nullptr;
    CK(c, cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    CK(c, cudaEventCreate(&e0)); CK(c, cudaEventCreate(&e1));
    const int iters = 1 << 14, blocks = c->num_sms * 8, threads = 256;
    for (int mode = 0; mode < 6; ++mode) {
        double best = 0;
        for (int rep = 0; rep < 4; ++rep) {
            CK(c, cudaEventRecord(e0, c->stream));
            if (mode == 0) k_fp64_rate<0><<<blocks, threads, 0, c->stream>>>(d, iters, 1.0 + rep);
            else if (mode == 1) k_fp64_rate<1><<<blocks, threads, 0, c->stream>>>(d, iters, 1.0 + rep);
            else if (mode == 2) k_fp64_rate<2><<<blocks, threads, 0, c->stream>>>(d, iters, 1.0 + rep);
            else if (mode == 3) k_fp64_mix<1><<<blocks, threads, 0, c->stream>>>(d, iters, 1.0 + rep);
            else if (mode == 4) k_fp64_mix<2><<<blocks, threads, 0, c->stream>>>(d, iters, 1.0 + rep);
            else k_fp64_mix<3><<<blocks, threads, 0, c->stream>>>(d, iters, 1.0 + rep);
            CK(c, cudaEventRecord(e1, c->stream));
            CK(c, cudaEventSynchronize(e1));
            c->launches++;
            float ms = 0;
            CK(c, cudaEventElapsedTime(&ms, e0, e1));
            const double inst = 8.0 * (double)iters * blocks * threads;
            if (rep > 0) best = std::max(best, inst / (ms * 1e-3) / 1e12);
        }
        out[mode] = best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return CCK_OK;
}
