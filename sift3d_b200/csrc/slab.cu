// slab.cu -- Z-slab tiling of ONE volume over several GPUs (SURVEY.md section 8e,
// BASELINE.json configs[4]).  The reference is single-process/host-memory and has no
// counterpart; what must be reproduced is its RESULT on the whole volume: the same pyramid
// bits, the same keypoints in the same (o, s, z, y, x) order, the same descriptors.
//
// Partition.  z is the slowest axis of the reference layout (SIFT3D_IM_GET_IDX,
// immacros.h:58-59), so the planes [a, b) a rank owns are one contiguous HBM range and a halo
// plane is one contiguous nx*ny*4-byte message.  Octave o+1 is the 2x decimation of octave o
// (im_downsample_2x, imutil.c:1742-1768: dst[z] = src[2z]), so a rank owns there exactly the
// planes whose source plane it owns: [ceil(a/2), ceil(b/2)) -- no alignment requirement on the
// split.  Every level buffer holds the owned planes plus `halo` planes either side (clipped to
// the volume), so local plane 0 / nz-1 are either true volume ends (where the reference's
// mirror rules, imutil.c:2365-2387, apply) or beyond the reach of every consumer.
//
// Exchanges (the only communication on the path):
//   * max|image| for im_scale (imutil.c:1983) and max|DoG| per level (sift.c:1161-1169):
//     all-reduce(max) of the u32 bit patterns (non-negative floats order like their bits);
//   * after every Gaussian level is computed on the owned planes: its halo planes, as many as
//     the level's consumers read (next blur's z reach, +-1 for DoG/extrema, the orientation
//     and descriptor windows for keypoint levels), from whichever ranks own them.
// Two transports behind one interface: NCCL send/recv groups (one process per GPU, the
// library is dlopen'ed so single-GPU users need no NCCL), and an in-process transport (one
// host thread per rank, cudaMemcpyPeerAsync + events) used for one-process multi-GPU runs
// and to test the tiling logic on a single GPU.
#include "common.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>

struct Xfer {
    int peer;
    void *ptr;
    size_t bytes;
};

struct s3d_comm {
    int rank = 0, nranks = 1;
    std::string err;
    virtual ~s3d_comm() {}
    virtual int allreduce_max_u32(unsigned *dev, int n, cudaStream_t st, int device) = 0;
    virtual int exchange(const std::vector<Xfer> &sends, const std::vector<Xfer> &recvs,
                         cudaStream_t st, int device) = 0;
};

namespace {

int comm_fail(s3d_comm *c, const std::string &msg)
{
    c->err = msg;
    fprintf(stderr, "sift3d_cuda comm (rank %d/%d): %s\n", c->rank, c->nranks, msg.c_str());
    return -1;
}

// ------------------------------------------------------------------ in-process transport
struct LocalWorld {
    int n = 0;
    std::mutex m;
    std::condition_variable cv;
    int waiting = 0;
    unsigned long gen = 0;
    bool broken = false;
    struct Msg {
        const void *ptr;
        size_t bytes;
        int dev;
        cudaEvent_t ready;
    };
    std::vector<std::deque<Msg>> box;  // [src * n + dst]
    std::vector<cudaEvent_t> done;     // per rank: receives of the current exchange finished
    std::vector<std::vector<unsigned>> red;

    bool barrier()
    {  // false on timeout (a rank failed before reaching it) -- never hang a test forever
        std::unique_lock<std::mutex> lk(m);
        if (broken) return false;
        const unsigned long g = gen;
        if (++waiting == n) {
            waiting = 0;
            gen++;
            cv.notify_all();
            return true;
        }
        if (!cv.wait_for(lk, std::chrono::seconds(300), [&] { return gen != g || broken; })) {
            broken = true;
            cv.notify_all();
            return false;
        }
        return !broken;
    }
};

struct LocalComm : s3d_comm {
    LocalWorld *w = nullptr;

    int allreduce_max_u32(unsigned *dev, int n, cudaStream_t st, int) override
    {
        std::vector<unsigned> &mine = w->red[rank];
        mine.resize(n);
        if (cudaMemcpyAsync(mine.data(), dev, n * sizeof(unsigned), cudaMemcpyDeviceToHost, st) !=
                cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return comm_fail(this, "allreduce: D2H failed");
        if (!w->barrier()) return comm_fail(this, "allreduce: barrier timeout");
        std::vector<unsigned> mx(n, 0u);
        for (int r = 0; r < nranks; r++) {
            if ((int)w->red[r].size() != n) return comm_fail(this, "allreduce: size mismatch");
            for (int i = 0; i < n; i++) mx[i] = std::max(mx[i], w->red[r][i]);
        }
        if (!w->barrier()) return comm_fail(this, "allreduce: barrier timeout");
        if (cudaMemcpyAsync(dev, mx.data(), n * sizeof(unsigned), cudaMemcpyHostToDevice, st) !=
                cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            return comm_fail(this, "allreduce: H2D failed");
        return 0;
    }

    int exchange(const std::vector<Xfer> &sends, const std::vector<Xfer> &recvs, cudaStream_t st,
                 int device) override
    {
        std::vector<cudaEvent_t> mine;
        for (const Xfer &s : sends) {
            cudaEvent_t ev;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventRecord(ev, st) != cudaSuccess)
                return comm_fail(this, "exchange: event");
            mine.push_back(ev);
            std::lock_guard<std::mutex> lk(w->m);
            w->box[(size_t)rank * nranks + s.peer].push_back({s.ptr, s.bytes, device, ev});
        }
        if (!w->barrier()) return comm_fail(this, "exchange: barrier timeout");
        for (const Xfer &r : recvs) {
            LocalWorld::Msg msg;
            {
                std::lock_guard<std::mutex> lk(w->m);
                auto &q = w->box[(size_t)r.peer * nranks + rank];
                if (q.empty()) return comm_fail(this, "exchange: missing message");
                msg = q.front();
                q.pop_front();
            }
            if (msg.bytes != r.bytes) return comm_fail(this, "exchange: size mismatch");
            cudaError_t ce = cudaStreamWaitEvent(st, msg.ready, 0);
            if (ce == cudaSuccess)
                ce = msg.dev == device
                         ? cudaMemcpyAsync(r.ptr, msg.ptr, r.bytes, cudaMemcpyDeviceToDevice, st)
                         : cudaMemcpyPeerAsync(r.ptr, device, msg.ptr, msg.dev, r.bytes, st);
            if (ce != cudaSuccess)
                return comm_fail(this, std::string("exchange: copy: ") + cudaGetErrorString(ce));
        }
        if (cudaEventRecord(w->done[rank], st) != cudaSuccess)
            return comm_fail(this, "exchange: done event");
        if (!w->barrier()) return comm_fail(this, "exchange: barrier timeout");
        // a sender must not overwrite what it sent before the receiver's copy has run
        for (const Xfer &s : sends)
            if (cudaStreamWaitEvent(st, w->done[s.peer], 0) != cudaSuccess)
                return comm_fail(this, "exchange: wait");
        if (!w->barrier()) return comm_fail(this, "exchange: barrier timeout");
        for (cudaEvent_t ev : mine) cudaEventDestroy(ev);
        return 0;
    }
};

// ------------------------------------------------------------------ NCCL transport
// Minimal declarations of the stable NCCL C ABI (nccl.h is not needed to build).
typedef struct {
    char internal[128];
} nccl_uid;
typedef void *nccl_comm_t;
enum { NCCL_CHAR = 0, NCCL_UINT32 = 3 };
enum { NCCL_MAX = 2 };

struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("SIFT3D_NCCL_LIB");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            if (!nm) continue;
            api.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.h) break;
        }
        if (!api.h) {
            api.err = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
            return;
        }
#define S3D_NCCL_SYM(field, name)                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.h, name)); \
    if (!api.field) api.err = std::string("missing symbol ") + name;
        S3D_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        S3D_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        S3D_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        S3D_NCCL_SYM(AllReduce, "ncclAllReduce")
        S3D_NCCL_SYM(Send, "ncclSend")
        S3D_NCCL_SYM(Recv, "ncclRecv")
        S3D_NCCL_SYM(GroupStart, "ncclGroupStart")
        S3D_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        S3D_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef S3D_NCCL_SYM
    });
    return &api;
}

struct NcclComm : s3d_comm {
    nccl_comm_t comm = nullptr;
    NcclApi *api = nullptr;

    ~NcclComm() override
    {
        if (comm) api->CommDestroy(comm);
    }
    int check(int rc, const char *what)
    {
        if (rc == 0) return 0;
        return comm_fail(this, std::string(what) + ": " + api->GetErrorString(rc));
    }
    int allreduce_max_u32(unsigned *dev, int n, cudaStream_t st, int) override
    {
        return check(api->AllReduce(dev, dev, (size_t)n, NCCL_UINT32, NCCL_MAX, comm, st),
                     "ncclAllReduce");
    }
    int exchange(const std::vector<Xfer> &sends, const std::vector<Xfer> &recvs, cudaStream_t st,
                 int) override
    {
        if (sends.empty() && recvs.empty()) return 0;
        if (check(api->GroupStart(), "ncclGroupStart")) return -1;
        int rc = 0;
        for (const Xfer &s : sends)
            if (!rc) rc = api->Send(s.ptr, s.bytes, NCCL_CHAR, s.peer, comm, st);
        for (const Xfer &r : recvs)
            if (!rc) rc = api->Recv(r.ptr, r.bytes, NCCL_CHAR, r.peer, comm, st);
        const int rc2 = api->GroupEnd();
        return check(rc ? rc : rc2, "ncclSend/ncclRecv group");
    }
};

// ------------------------------------------------------------------ geometry
inline int ceil_div2(int v) { return (v + 1) / 2; }

// halo planes a keypoint window on Gaussian level `g` reads beyond its centre plane:
// the descriptor sphere (extract_descrip, sift.c:1845-1850: radius 2 * 7.07 * sd, in
// physical units) is the widest consumer; + 1 for the central-difference gradient, + 1 for
// the ceil() of the loop bounds (IM_LOOP_SPHERE_START, sift.c:96-119).
int window_need(const s3d_geom &g)
{
    const float sigma = (float)(g.scale * 7.071067812);
    const float rad = (float)(2.0 * (double)sigma);
    return (int)std::ceil((double)rad / g.uz) + 2;
}

}  // namespace

// ------------------------------------------------------------------ slab pipeline (internal)
// Halo transfers of one level: pure host logic, shared by the pipeline and the CPU tests.
// own = [rank][octave] -> (own0, own1).  Emits {kind (0 send / 1 recv), peer, z0, z1} in the
// order both sides of a pair agree on: per peer, the receiver's low interval, then its high one.
struct HaloXfer {
    int kind, peer, z0, z1;
};

static void halo_plan(int nranks, int noct, const int *own, int o, int NZ, int h, int me,
                      std::vector<HaloXfer> &out)
{
    const int a = own[2 * ((size_t)me * noct + o)], b = own[2 * ((size_t)me * noct + o) + 1];
    auto add = [&](int kind, int peer, int z0, int z1) {
        if (z1 > z0) out.push_back({kind, peer, z0, z1});
    };
    for (int p = 0; p < nranks; p++) {
        if (p == me) continue;
        const int pa = own[2 * ((size_t)p * noct + o)], pb = own[2 * ((size_t)p * noct + o) + 1];
        if (pb <= pa || b <= a) continue;
        // what I receive from p: my halo intervals (low, then high) intersected with p's planes
        add(1, p, std::max(std::max(a - h, 0), pa), std::min(a, pb));
        add(1, p, std::max(b, pa), std::min(std::min(b + h, NZ), pb));
        // what I send to p: p's halo intervals (same order) intersected with my planes
        add(0, p, std::max(std::max(pa - h, 0), a), std::min(pa, b));
        add(0, p, std::max(pb, a), std::min(std::min(pb + h, NZ), b));
    }
}

// CUDA events around a communication step (option "slab_timing"): pairs are summed after the
// pyramid is built.  The interval also contains the wait for the slower neighbour.
static void slab_mark(s3d_engine *e, int kind, bool begin)
{
    if (!e->opt_slab_timing) return;
    cudaEvent_t ev = nullptr;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, e->stream);
    e->slab_ev.push_back(ev);
    if (begin) e->slab_ev_kind.push_back(kind);
}

static int slab_exchange(s3d_engine *e, float *base, int o, int h)
{
    S3dRange nvtx_range("s3d:halo_exchange");
    s3d_comm *c = e->comm;
    const SlabOct &S = e->slab[o];
    const s3d_geom &g0 = e->slab_g[(size_t)o * e->nlev_g];
    const size_t plane = (size_t)g0.nx * g0.ny;
    std::vector<HaloXfer> plan;
    halo_plan(c->nranks, e->noct, e->slab_own.data(), o, S.NZ, h, c->rank, plan);
    std::vector<Xfer> sends, recvs;
    for (const HaloXfer &x : plan)
        (x.kind ? recvs : sends)
            .push_back({x.peer, base + (size_t)(x.z0 - S.lo) * plane,
                        (size_t)(x.z1 - x.z0) * plane * sizeof(float)});
    for (const Xfer &x : sends) e->slab_sent += (double)x.bytes;
    for (const Xfer &x : recvs) e->slab_recv += (double)x.bytes;
    if (!sends.empty() || !recvs.empty()) e->slab_nxchg++;
    slab_mark(e, 0, true);
    if (c->exchange(sends, recvs, e->stream, e->device)) {
        e->err = c->err;
        return -1;
    }
    slab_mark(e, 0, false);
    return 0;
}

int s3d_slab_build_pyramid(s3d_engine *e)
{
    s3d_comm *c = e->comm;
    if (!c) return s3d_fail(e, "slab: no communicator", cudaSuccess, __FILE__, __LINE__);
    const int nlg = e->nlev_g, nld = e->nlev_d;
    const SlabOct &S0 = e->slab[0];
    const s3d_geom &G0 = e->slab_g[0];
    const size_t plane0 = (size_t)G0.nx * G0.ny;
    e->slab_sent = e->slab_recv = e->slab_xchg_ms = e->slab_allreduce_ms = 0.0;
    e->slab_nxchg = 0;
    for (cudaEvent_t ev : e->slab_ev) cudaEventDestroy(ev);
    e->slab_ev.clear();
    e->slab_ev_kind.clear();
    S3D_CUDA(e, cudaMemsetAsync(e->d_scalars, 0, e->n_scalars * sizeof(unsigned), e->stream));
    // im_scale (imutil.c:1977-1991): max over the owned planes, then over the ranks
    float *own_im = e->im + (size_t)(S0.own0 - S0.lo) * plane0;
    const size_t n_own = (size_t)(S0.own1 - S0.own0) * plane0;
    if (n_own && s3d_k_max_abs(e, own_im, n_own, e->d_scalars)) return -1;
    slab_mark(e, 1, true);
    if (c->allreduce_max_u32(e->d_scalars, 1, e->stream, e->device)) {
        e->err = c->err;
        return -1;
    }
    slab_mark(e, 1, false);
    if (n_own && s3d_k_scale(e, own_im, own_im, n_own, e->d_scalars)) return -1;
    float uf[3];
    auto level_uf = [&](const s3d_geom &g) {
        uf[0] = (float)(1.0 / g.ux);
        uf[1] = (float)(1.0 / g.uy);
        uf[2] = (float)(1.0 / g.uz);
    };
    // build_gpyr (sift.c:989-1050)
    level_uf(G0);
    if (slab_exchange(e, e->im, 0, s3d_blur_z_reach(e->first_taps, uf[2]))) return -1;
    if (s3d_k_blur_zrange(e, e->im, e->g[0].d, G0.nx, G0.ny, S0.hi - S0.lo, e->first_taps, uf,
                          S0.own0 - S0.lo, S0.own1 - S0.lo, S0.lo, S0.NZ))
        return -1;
    for (int o = 0; o < e->noct; o++) {
        const SlabOct &S = e->slab[o];
        const s3d_geom &G = e->slab_g[(size_t)o * nlg];
        level_uf(G);
        for (int s = 0; s < nlg; s++) {
            LevelDev &cur = e->g[(size_t)o * nlg + s];
            if (slab_exchange(e, cur.d, o, e->slab_need[(size_t)o * nlg + s])) return -1;
            if (s == nlg - 1) break;
            LevelDev &dst = e->g[(size_t)o * nlg + s + 1];
            if (s3d_k_blur_zrange(e, cur.d, dst.d, G.nx, G.ny, S.hi - S.lo, e->oct_taps[s], uf,
                                  S.own0 - S.lo, S.own1 - S.lo, S.lo, S.NZ))
                return -1;
        }
        if (o != e->noct - 1) {  // im_downsample_2x of level max(s_end-2, first) (sift.c:1029-1041)
            const int ds = std::max(nlg - 2 - 2, -1);
            const LevelDev &src = e->g[(size_t)o * nlg + ds + 1];
            LevelDev &dst = e->g[(size_t)(o + 1) * nlg];
            const SlabOct &D = e->slab[o + 1];
            const s3d_geom &GD = e->slab_g[(size_t)(o + 1) * nlg];
            if (D.own1 > D.own0) {
                const float *sp = src.d + (size_t)(2 * D.own0 - S.lo) * G.nx * G.ny;
                float *dp = dst.d + (size_t)(D.own0 - D.lo) * GD.nx * GD.ny;
                if (s3d_k_decimate(e, sp, G.nx, G.ny, 0, dp, GD.nx, GD.ny, D.own1 - D.own0)) return -1;
            }
        }
    }
    // build_dog (sift.c:1052-1071) on the owned planes +-1, with the per-level max|DoG|
    for (int o = 0; o < e->noct; o++) {
        const SlabOct &S = e->slab[o];
        if (S.own1 <= S.own0) continue;
        const s3d_geom &G = e->slab_g[(size_t)o * nlg];
        const int z0 = std::max(S.lo, S.own0 - 1), z1 = std::min(S.hi, S.own1 + 1);
        const size_t off = (size_t)(z0 - S.lo) * G.nx * G.ny, n = (size_t)(z1 - z0) * G.nx * G.ny;
        for (int s = 0; s < nld; s++)
            if (s3d_k_dog(e, e->g[(size_t)o * nlg + s].d + off, e->g[(size_t)o * nlg + s + 1].d + off,
                          e->dog[(size_t)o * nld + s].d + off, n, e->d_scalars + 1 + (size_t)o * nld + s))
                return -1;
    }
    slab_mark(e, 1, true);
    if (c->allreduce_max_u32(e->d_scalars, e->n_scalars, e->stream, e->device)) {
        e->err = c->err;
        return -1;
    }
    slab_mark(e, 1, false);
    return 0;
}

int s3d_slab_stats(s3d_engine *e, double out[5])
{
    if (e->slab.empty()) return -1;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    if (!e->slab_ev.empty()) {  // timed run: sum the event pairs once
        cudaStreamSynchronize(e->stream);
        e->slab_xchg_ms = e->slab_allreduce_ms = 0.0;
        for (size_t i = 0; 2 * i + 1 < e->slab_ev.size() && i < e->slab_ev_kind.size(); i++) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, e->slab_ev[2 * i], e->slab_ev[2 * i + 1]) == cudaSuccess)
                (e->slab_ev_kind[i] ? e->slab_allreduce_ms : e->slab_xchg_ms) += ms;
        }
    }
    out[0] = e->slab_sent;
    out[1] = e->slab_recv;
    out[2] = (double)e->slab_nxchg;
    out[3] = e->opt_slab_timing ? e->slab_xchg_ms : -1.0;
    out[4] = e->opt_slab_timing ? e->slab_allreduce_ms : -1.0;
    if (prev >= 0) cudaSetDevice(prev);
    return 0;
}

int s3d_slab_extrema_octave(s3d_engine *e, int o, double peak_thresh)
{  // detect_extrema scans z = 1 .. NZ-2 (sift.c:1182-1190); this rank scans its share
    const SlabOct &S = e->slab[o];
    const int z0 = std::max(S.own0, 1), z1 = std::min(S.own1, S.NZ - 1);
    if (z1 <= z0) return 0;
    return s3d_k_extrema_range(e, o, peak_thresh, z0 - 1 - S.lo, z1 - z0 + 2, z0 - 1);
}

// ------------------------------------------------------------------ C ABI
extern "C" {

void *s3d_local_world_create(int nranks)
{
    if (nranks < 1) return nullptr;
    LocalWorld *w = new LocalWorld();
    w->n = nranks;
    w->box.resize((size_t)nranks * nranks);
    w->done.assign(nranks, nullptr);
    w->red.resize(nranks);
    return w;
}

void s3d_local_world_destroy(void *world)
{
    LocalWorld *w = static_cast<LocalWorld *>(world);
    if (!w) return;
    for (cudaEvent_t ev : w->done)
        if (ev) cudaEventDestroy(ev);
    delete w;
}

int s3d_comm_create_local(s3d_comm **out, void *world, int rank, int device)
{
    *out = nullptr;
    LocalWorld *w = static_cast<LocalWorld *>(world);
    if (!w || rank < 0 || rank >= w->n) return -1;
    int prev = -1;
    cudaGetDevice(&prev);
    if (device >= 0) cudaSetDevice(device);
    cudaEvent_t ev = nullptr;
    const cudaError_t ce = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (device >= 0 && prev >= 0) cudaSetDevice(prev);
    if (ce != cudaSuccess) {
        fprintf(stderr, "sift3d_cuda: s3d_comm_create_local: %s\n", cudaGetErrorString(ce));
        return -1;
    }
    {
        std::lock_guard<std::mutex> lk(w->m);
        w->done[rank] = ev;
    }
    LocalComm *c = new LocalComm();
    c->rank = rank;
    c->nranks = w->n;
    c->w = w;
    *out = c;
    return 0;
}

int s3d_nccl_unique_id(unsigned char id[128])
{
    NcclApi *api = nccl_api();
    if (!api->err.empty()) {
        fprintf(stderr, "sift3d_cuda: NCCL unavailable: %s\n", api->err.c_str());
        return -1;
    }
    nccl_uid u;
    const int rc = api->GetUniqueId(&u);
    if (rc) {
        fprintf(stderr, "sift3d_cuda: ncclGetUniqueId: %s\n", api->GetErrorString(rc));
        return -1;
    }
    memcpy(id, u.internal, 128);
    return 0;
}

int s3d_comm_create_nccl(s3d_comm **out, int rank, int nranks, const unsigned char id[128],
                         int device)
{
    *out = nullptr;
    NcclApi *api = nccl_api();
    if (!api->err.empty()) {
        fprintf(stderr, "sift3d_cuda: NCCL unavailable: %s\n", api->err.c_str());
        return -1;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return -1;
    nccl_uid u;
    memcpy(u.internal, id, 128);
    NcclComm *c = new NcclComm();
    c->rank = rank;
    c->nranks = nranks;
    c->api = api;
    const int rc = api->CommInitRank(&c->comm, nranks, u, rank);
    if (device >= 0 && prev >= 0) cudaSetDevice(prev);
    if (rc) {
        fprintf(stderr, "sift3d_cuda: ncclCommInitRank: %s\n", api->GetErrorString(rc));
        c->comm = nullptr;
        delete c;
        return -1;
    }
    *out = c;
    return 0;
}

void s3d_comm_destroy(s3d_comm *c) { delete c; }
int s3d_comm_rank(const s3d_comm *c) { return c->rank; }
int s3d_comm_size(const s3d_comm *c) { return c->nranks; }
const char *s3d_comm_error(const s3d_comm *c) { return c->err.c_str(); }

int s3d_comm_allreduce_max_u32(s3d_comm *c, s3d_engine *e, unsigned *dev, int n)
{  // exposed for tests of the transports
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    const int rc = c->allreduce_max_u32(dev, n, e->stream, e->device);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int s3d_slab_plan(int nranks, int num_octaves, int nz0, const int *zsplit, int *own /* [nranks][num_octaves][2] */)
{  // pure host logic (also used by the CPU tests): the planes every rank owns per octave
    if (nranks < 1 || num_octaves < 1 || zsplit[0] != 0 || zsplit[nranks] != nz0) return -1;
    for (int r = 0; r < nranks; r++) {
        if (zsplit[r + 1] < zsplit[r]) return -1;
        int a = zsplit[r], b = zsplit[r + 1], NZ = nz0;
        for (int o = 0; o < num_octaves; o++) {
            own[2 * ((size_t)r * num_octaves + o)] = std::min(a, NZ);
            own[2 * ((size_t)r * num_octaves + o) + 1] = std::min(b, NZ);
            a = ceil_div2(a);
            b = ceil_div2(b);
            NZ /= 2;
        }
    }
    return 0;
}

int s3d_slab_halo_plan(int nranks, int num_octaves, const int *own, int o, int NZ, int h, int rank,
                       int *out, int cap)
{  // host logic, for tests: rows of {kind, peer, z0, z1}; returns the row count (or -1)
    std::vector<HaloXfer> plan;
    halo_plan(nranks, num_octaves, own, o, NZ, h, rank, plan);
    if ((int)plan.size() > cap) return -1;
    for (size_t i = 0; i < plan.size(); i++) {
        out[4 * i] = plan[i].kind;
        out[4 * i + 1] = plan[i].peer;
        out[4 * i + 2] = plan[i].z0;
        out[4 * i + 3] = plan[i].z1;
    }
    return (int)plan.size();
}

static int slab_pyramid_resize_impl(s3d_engine *e, s3d_comm *comm, int num_octaves, int num_kp_levels,
                                    const s3d_geom *gpyr, const s3d_geom *dog, const int *zsplit);

int s3d_slab_pyramid_resize(s3d_engine *e, s3d_comm *comm, int num_octaves, int num_kp_levels,
                            const s3d_geom *gpyr, const s3d_geom *dog, const int *zsplit)
{
    // a failure half way (cudaMalloc) must not leave geometry that the `same` fast path of the
    // next call would accept with null level pointers: drop everything
    const int rc = slab_pyramid_resize_impl(e, comm, num_octaves, num_kp_levels, gpyr, dog, zsplit);
    if (rc) {
        const std::string msg = e->err;
        s3d_free_pyramid(e);
        e->err = msg;
    }
    return rc;
}

static int slab_pyramid_resize_impl(s3d_engine *e, s3d_comm *comm, int num_octaves, int num_kp_levels,
                                    const s3d_geom *gpyr, const s3d_geom *dog, const int *zsplit)
{
    int prev = -1;
    cudaGetDevice(&prev);
    struct Restore {
        int p;
        ~Restore()
        {
            if (p >= 0) cudaSetDevice(p);
        }
    } restore{prev};
    cudaSetDevice(e->device);
    if (!comm || num_octaves < 1 || (int)e->oct_taps.size() != num_kp_levels + 2)
        return s3d_fail(e, "s3d_slab_pyramid_resize: set the filters (s3d_pyramid_filters) first",
                        cudaSuccess, __FILE__, __LINE__);
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    const int nlg = num_kp_levels + 3, nld = num_kp_levels + 2;
    const int R = comm->nranks, me = comm->rank;
    {   // same tiling as last time: keep every allocation, refresh units / scales only
        bool same = !e->slab.empty() && e->comm == comm && e->noct == num_octaves &&
                    e->K == num_kp_levels && (int)e->zsplit.size() == R + 1 &&
                    std::equal(zsplit, zsplit + R + 1, e->zsplit.begin()) &&
                    e->slab_g.size() == (size_t)num_octaves * nlg;
        for (size_t i = 0; same && i < e->slab_g.size(); i++)
            same = e->slab_g[i].nx == gpyr[i].nx && e->slab_g[i].ny == gpyr[i].ny &&
                   e->slab_g[i].nz == gpyr[i].nz && e->slab_g[i].ux == gpyr[i].ux &&
                   e->slab_g[i].uy == gpyr[i].uy && e->slab_g[i].uz == gpyr[i].uz &&
                   e->slab_g[i].scale == gpyr[i].scale;
        if (same) return 0;
    }
    s3d_free_pyramid(e);
    e->comm = comm;
    e->noct = num_octaves;
    e->K = num_kp_levels;
    e->nlev_g = nlg;
    e->nlev_d = nld;
    e->first_level = -1;
    e->zsplit.assign(zsplit, zsplit + R + 1);
    e->slab_g.assign(gpyr, gpyr + (size_t)num_octaves * nlg);
    e->slab_own.resize((size_t)2 * R * num_octaves);
    if (s3d_slab_plan(R, num_octaves, gpyr[0].nz, zsplit, e->slab_own.data()))
        return s3d_fail(e, "s3d_slab_pyramid_resize: bad z split", cudaSuccess, __FILE__, __LINE__);
    // halo every consumer of a level needs
    e->slab_need.assign((size_t)num_octaves * nlg, 1);
    int halo = 1;
    for (int o = 0; o < num_octaves; o++)
        for (int s = 0; s < nlg; s++) {
            const s3d_geom &g = gpyr[(size_t)o * nlg + s];
            int need = 1;  // DoG / extrema neighbours
            if (s + 1 < nlg) need = std::max(need, s3d_blur_z_reach(e->oct_taps[s], (float)(1.0 / g.uz)));
            if (s >= 1 && s <= num_kp_levels) need = std::max(need, window_need(g));  // levels 0..K-1
            e->slab_need[(size_t)o * nlg + s] = need;
            halo = std::max(halo, need);
        }
    halo = std::max(halo, s3d_blur_z_reach(e->first_taps, (float)(1.0 / gpyr[0].uz)));
    e->slab_halo = halo;
    e->slab.resize(num_octaves);
    for (int o = 0; o < num_octaves; o++) {
        SlabOct &S = e->slab[o];
        S.NZ = gpyr[(size_t)o * nlg].nz;
        S.own0 = e->slab_own[2 * ((size_t)me * num_octaves + o)];
        S.own1 = e->slab_own[2 * ((size_t)me * num_octaves + o) + 1];
        if (S.own1 > S.own0) {
            S.lo = std::max(0, S.own0 - halo);
            S.hi = std::min(S.NZ, S.own1 + halo);
        } else {
            S.lo = S.hi = S.own0;
        }
    }
    e->g.resize((size_t)num_octaves * nlg);
    e->dog.resize((size_t)num_octaves * nld);
    for (int o = 0; o < num_octaves; o++) {
        const SlabOct &S = e->slab[o];
        for (int s = 0; s < nlg + nld; s++) {
            LevelDev &l = s < nlg ? e->g[(size_t)o * nlg + s] : e->dog[(size_t)o * nld + (s - nlg)];
            l.g = gpyr[(size_t)o * nlg + (s < nlg ? s : s - nlg)];
            if (s >= nlg) l.g.scale = dog[(size_t)o * nld + (s - nlg)].scale;
            l.g.nz = S.hi - S.lo;
            l.d = nullptr;
            if (l.n()) {
                S3D_CUDA(e, cudaMalloc(&l.d, l.n() * sizeof(float)));
                // halo planes no consumer needs are never written; keep them finite
                S3D_CUDA(e, cudaMemsetAsync(l.d, 0, l.n() * sizeof(float), e->stream));
            }
        }
    }
    {   // the scaled input copy: geometry of octave 0
        const SlabOct &S = e->slab[0];
        const size_t n = (size_t)gpyr[0].nx * gpyr[0].ny * std::max(S.hi - S.lo, 1);
        if (n > e->im_cap) {
            if (e->im) cudaFree(e->im);
            e->im = nullptr;
            e->im_cap = 0;
            S3D_CUDA(e, cudaMalloc(&e->im, n * sizeof(float)));
            e->im_cap = n;
        }
        S3D_CUDA(e, cudaMemsetAsync(e->im, 0, n * sizeof(float), e->stream));
        e->im_nx = gpyr[0].nx;
        e->im_ny = gpyr[0].ny;
        e->im_nz = S.hi - S.lo;
    }
    const int L = num_octaves * nlg;
    S3D_CUDA(e, cudaMalloc(&e->d_level_ptrs, L * sizeof(float *)));
    S3D_CUDA(e, cudaMalloc(&e->d_level_dims, 3 * L * sizeof(int)));
    S3D_CUDA(e, cudaMalloc(&e->d_level_units, 3 * L * sizeof(float)));
    S3D_CUDA(e, cudaMalloc(&e->d_level_scales, L * sizeof(double)));
    S3D_CUDA(e, cudaMalloc(&e->d_level_zoff, L * sizeof(int)));
    e->n_scalars = 1 + num_octaves * nld;
    S3D_CUDA(e, cudaMalloc(&e->d_scalars, e->n_scalars * sizeof(unsigned)));
    std::vector<float *> ptrs(L);
    std::vector<int> dims(3 * L), zoff(L);
    std::vector<float> units(3 * L);
    std::vector<double> scales(L);
    for (int i = 0; i < L; i++) {
        ptrs[i] = e->g[i].d;
        dims[3 * i] = e->g[i].g.nx;
        dims[3 * i + 1] = e->g[i].g.ny;
        dims[3 * i + 2] = e->g[i].g.nz;
        units[3 * i] = (float)gpyr[i].ux;
        units[3 * i + 1] = (float)gpyr[i].uy;
        units[3 * i + 2] = (float)gpyr[i].uz;
        scales[i] = gpyr[i].scale;
        zoff[i] = e->slab[i / nlg].lo;
    }
    S3D_CUDA(e, cudaMemcpy(e->d_level_ptrs, ptrs.data(), L * sizeof(float *), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_dims, dims.data(), 3 * L * sizeof(int), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_units, units.data(), 3 * L * sizeof(float), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_scales, scales.data(), L * sizeof(double), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaMemcpy(e->d_level_zoff, zoff.data(), L * sizeof(int), cudaMemcpyHostToDevice));
    S3D_CUDA(e, cudaStreamSynchronize(e->stream));
    return 0;
}

int s3d_slab_info(const s3d_engine *e, int o, int info[6])
{
    if (e->slab.empty() || o < 0 || o >= (int)e->slab.size()) return -1;
    const SlabOct &S = e->slab[o];
    info[0] = S.own0;
    info[1] = S.own1;
    info[2] = S.lo;
    info[3] = S.hi;
    info[4] = S.NZ;
    info[5] = e->slab_halo;
    return 0;
}

int s3d_slab_image_upload(s3d_engine *e, const float *host, size_t xs, size_t ys, size_t zs)
{  // the rank's OWNED planes of the input (im_copy_data, imutil.c:1895), nx * ny * (own1 - own0)
    if (e->slab.empty())
        return s3d_fail(e, "s3d_slab_image_upload: not in slab mode", cudaSuccess, __FILE__, __LINE__);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    const SlabOct &S = e->slab[0];
    const int nx = e->im_nx, ny = e->im_ny, nz = S.own1 - S.own0;
    int rc = 0;
    if (nz > 0) {
        float *dst = e->im + (size_t)(S.own0 - S.lo) * nx * ny;
        if (xs == 1 && ys == (size_t)nx && zs == (size_t)nx * ny) {
            if (cudaMemcpyAsync(dst, host, (size_t)nx * ny * nz * sizeof(float), cudaMemcpyHostToDevice,
                                e->stream) != cudaSuccess)
                rc = s3d_fail(e, "slab upload", cudaGetLastError(), __FILE__, __LINE__);
        } else {
            std::vector<float> tmp((size_t)nx * ny * nz);
            for (int z = 0; z < nz; z++)
                for (int y = 0; y < ny; y++)
                    for (int x = 0; x < nx; x++)
                        tmp[x + (size_t)nx * (y + (size_t)ny * z)] = host[x * xs + y * ys + z * zs];
            if (cudaMemcpyAsync(dst, tmp.data(), tmp.size() * sizeof(float), cudaMemcpyHostToDevice,
                                e->stream) != cudaSuccess ||
                cudaStreamSynchronize(e->stream) != cudaSuccess)
                rc = s3d_fail(e, "slab upload", cudaGetLastError(), __FILE__, __LINE__);
        }
    }
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int s3d_slab_image_from_device(s3d_engine *e, const float *dev)
{
    if (e->slab.empty())
        return s3d_fail(e, "s3d_slab_image_from_device: not in slab mode", cudaSuccess, __FILE__, __LINE__);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(e->device);
    const SlabOct &S = e->slab[0];
    const size_t plane = (size_t)e->im_nx * e->im_ny;
    int rc = 0;
    if (S.own1 > S.own0 &&
        cudaMemcpyAsync(e->im + (size_t)(S.own0 - S.lo) * plane, dev,
                        plane * (S.own1 - S.own0) * sizeof(float), cudaMemcpyDeviceToDevice,
                        e->stream) != cudaSuccess)
        rc = s3d_fail(e, "slab device copy", cudaGetLastError(), __FILE__, __LINE__);
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

}  // extern "C"
