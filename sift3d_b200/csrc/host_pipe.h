// host_pipe.h -- host threads for the copies between PAGEABLE host memory and the device
// (engine.cu): the caller's Image goes up (im_copy_data, imutil.c:1895) and results come down
// into the caller's malloc memory (the reference's ownership contract), so every large transfer
// is staged through pinned memory.  Three ways, option "copy_pipe" of include/sift3d_cuda.h:
//   0  HostTeam::copy    one chunk at a time, split over the team and the caller
//   1  PipeJob           workers move whole chunks through a ring of slots, the caller issues the
//                        DMAs and polls (the default when the process has the host to itself)
//   2  Pipe2Job          everybody owns a stream and two slots and runs its chunks alone (the
//                        default when ranks share the host)
// Plain C++ (no CUDA): tools/pipe_host_check.cpp runs both pipelines against a mock DMA engine
// on the CPU (tests/test_host_pipe.py).
#pragma once

#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace {
static inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}

class HostTeam {
public:
    static HostTeam &get()
    {
        static HostTeam t;
        return t;
    }
    unsigned workers() const { return (unsigned)workers_.size(); }
    unsigned ranks_on_host() const { return share_; }  // $LOCAL_WORLD_SIZE (1 when not set)
    // Runs fn(t) on the workers t = first .. workers() - 1 while the caller runs main_fn();
    // returns when all of them are done.  One job at a time (engines share the team).
    void run(unsigned first, const std::function<void(unsigned)> &job, const std::function<void()> &main_fn)
    {
        std::unique_lock<std::mutex> lk(call_mu_);
        {
            std::lock_guard<std::mutex> g(mu_);
            job_ = &job;
            first_ = first;
            pending_ = (int)workers_.size();
            gen_++;
        }
        cv_.notify_all();
        main_fn();
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [&] { return pending_ == 0; });
        job_ = nullptr;
    }
    // memcpy split into page-aligned shares; the caller takes share 0, worker 0 sits it out
    // (it only serves the chunk pipeline below)
    void copy(char *d, const char *src, size_t len)
    {
        const unsigned n = (unsigned)workers_.size();  // shares: caller + workers 1 .. n - 1
        const size_t part = ((len + n - 1) / n + 4095) & ~(size_t)4095;
        run(1,
            [=](unsigned t) {
                const size_t lo = (size_t)t * part;
                if (lo < len) memcpy(d + lo, src + lo, std::min(part, len - lo));
            },
            [=] { memcpy(d, src, std::min(part, len)); });
    }

private:
    HostTeam()
    {
        unsigned hw = std::thread::hardware_concurrency();
        if (hw < 1) hw = 1;
        unsigned share = 1;
        if (const char *v = getenv("LOCAL_WORLD_SIZE")) share = (unsigned)std::max(1, atoi(v));
        share_ = share;
        unsigned n = std::max(2u, std::min(8u, hw / share));  // measured on a 16-core host: 8 beats 16
        if (const char *v = getenv("S3D_COPY_THREADS")) n = (unsigned)std::max(1, atoi(v));
        for (unsigned t = 0; t < n; t++) workers_.emplace_back([this, t] { loop(t); });
    }
    ~HostTeam()
    {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
            gen_++;
        }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    void loop(unsigned t)
    {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> g(mu_);
            cv_.wait(g, [&] { return gen_ != seen; });
            seen = gen_;
            if (stop_) return;
            const std::function<void(unsigned)> *job = job_;
            const unsigned first = first_;
            g.unlock();
            if (t >= first) (*job)(t);
            g.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_;
    unsigned long long gen_ = 0;
    int pending_ = 0;
    bool stop_ = false;
    const std::function<void(unsigned)> *job_ = nullptr;
    unsigned first_ = 0;
    unsigned share_ = 1;
};

// The chunk pipeline between pageable host memory and a ring of pinned slots.  The parallel
// memcpy above splits ONE chunk over the team and meets at a condition variable per chunk: the
// DMA of chunk c and the copy of chunk c + 1 alternate, and the wake-ups cost 50-100 us each.
// Here every worker owns WHOLE chunks (claimed from an atomic counter) and the chunks of a
// transfer flow through the ring without any per-chunk rendezvous: the calling thread only
// issues the DMAs in order and polls their events, the workers only copy.
//   upload:   worker fills slot c % ns once chunk c - ns has left it (gate = chunks drained),
//             raises flag[c]; the caller issues the DMA of chunk c when flag[c] is up.
//   download: the caller issues the DMA of chunk c into slot c % ns once flag[c - ns] is up (a
//             worker has copied that chunk out), polls the events and publishes gate = chunks
//             arrived; a worker copies chunk c out when c < gate and raises flag[c].
// Small slots (a few MB) keep the ring inside the CPU's last-level cache, so the DMA engine's
// reads (upload) hit lines the workers have just written instead of going to DRAM and back.
struct PipeJob {
    int dir = 0;  // 0 upload (pageable -> slots), 1 download (slots -> pageable)
    char *host = nullptr;
    char *slots = nullptr;
    size_t bytes = 0, ch = 0, nch = 0, ns = 0;
    std::atomic<size_t> next{0}, gate{0}, ndone{0};
    std::atomic<int> abort{0};
    std::atomic<unsigned char> *flag = nullptr;
};

inline void pipe_worker(PipeJob &J)
{
    for (;;) {
        const size_t c = J.next.fetch_add(1, std::memory_order_relaxed);
        if (c >= J.nch) return;
        for (unsigned spins = 0;; spins++) {
            const size_t g = J.gate.load(std::memory_order_acquire);
            if (J.dir == 0 ? c < g + J.ns : c < g) break;
            if (J.abort.load(std::memory_order_relaxed)) return;
            if (spins > 256) sched_yield();
            else cpu_relax();
        }
        char *slot = J.slots + (c % J.ns) * J.ch;
        const size_t off = c * J.ch, len = std::min(J.ch, J.bytes - off);
        if (J.dir == 0) memcpy(slot, J.host + off, len);
        else memcpy(J.host + off, slot, len);
        J.flag[c].store(1, std::memory_order_release);
        J.ndone.fetch_add(1, std::memory_order_release);
    }
}

// The calling thread's side of a transfer: issue(c) queues the DMA of chunk c (in order) and
// returns false on error; poll(c) tells whether the DMA of chunk c has completed (1), is in
// flight (0) or failed (-1).  The workers of the team run pipe_worker meanwhile.  Returns false
// when issue / poll reported an error (the workers have left the job by then).
template <class Issue, class Poll>
inline bool s3d_pipe_run(PipeJob &J, Issue &&issue, Poll &&poll, unsigned first_worker = 0)
{
    bool ok = true;
    auto idle = [](unsigned &spins) {
        if (++spins > 256) sched_yield();
        else cpu_relax();
    };
    auto main_up = [&] {
        size_t issued = 0, drained = 0;
        unsigned spins = 0;
        while (issued < J.nch) {
            bool progress = false;
            if (drained < issued) {
                const int q = poll(drained);
                if (q < 0) {
                    ok = false;
                    break;
                }
                if (q > 0) {
                    drained++;
                    J.gate.store(drained, std::memory_order_release);
                    progress = true;
                }
            }
            if (J.flag[issued].load(std::memory_order_acquire)) {
                if (!issue(issued)) {
                    ok = false;
                    break;
                }
                issued++;
                progress = true;
            }
            if (progress) spins = 0;
            else idle(spins);
        }
        if (!ok) J.abort.store(1);
    };
    auto main_down = [&] {
        size_t issued = 0, arrived = 0;
        unsigned spins = 0;
        while (J.ndone.load(std::memory_order_acquire) < J.nch) {
            bool progress = false;
            if (issued < J.nch && (issued < J.ns || J.flag[issued - J.ns].load(std::memory_order_acquire))) {
                if (!issue(issued)) {
                    ok = false;
                    break;
                }
                issued++;
                progress = true;
            }
            if (arrived < issued) {
                const int q = poll(arrived);
                if (q < 0) {
                    ok = false;
                    break;
                }
                if (q > 0) {
                    arrived++;
                    J.gate.store(arrived, std::memory_order_release);
                    progress = true;
                }
            }
            if (progress) spins = 0;
            else idle(spins);
        }
        if (!ok) J.abort.store(1);
    };
    if (J.dir == 0) HostTeam::get().run(first_worker, [&](unsigned) { pipe_worker(J); }, main_up);
    else HostTeam::get().run(first_worker, [&](unsigned) { pipe_worker(J); }, main_down);
    return ok;
}
// ---- the pipeline without a polling thread (option copy_pipe = 2) -------------------------
// When several ranks share a host a rank has 2 ... 4 copy threads, and one that only issues and
// polls is a quarter or half of the copy capacity gone.  Here EVERY participant -- the workers
// and the caller -- owns two private pinned slots, its own stream and two events, and runs the
// whole life of its chunks itself: claim a chunk from the atomic counter, wait for the slot's
// previous DMA, copy, queue the DMA, record the event.  No shared state but the counter.
//   issue(p, c, k): queue the DMA of chunk c between participant p's slot k and the device on p's
//                   stream and record the slot's event; false on error
//   wait(p, k):     block until the DMA last recorded for p's slot k has completed; false on error
//   enter(p):       called once by participant p on its own thread before anything else
// The caller is participant workers(); its slots are the last two.  Returns false on error.  The
// DMAs of an upload may still be in flight on return: the caller joins the streams.
struct Pipe2Job {
    int dir = 0;
    char *host = nullptr;
    char *slots = nullptr;  // 2 * (workers + 1) slots of ch bytes
    size_t bytes = 0, ch = 0, nch = 0;
    std::atomic<size_t> next{0};
    std::atomic<int> fail{0};
};

template <class Enter, class Issue, class Wait>
inline void s3d_pipe2_participant(Pipe2Job &J, unsigned p, Enter &enter, Issue &issue, Wait &wait)
{
    if (!enter(p)) {
        J.fail.store(1);
        return;
    }
    char *slot[2] = {J.slots + (size_t)(2 * p) * J.ch, J.slots + (size_t)(2 * p + 1) * J.ch};
    auto len_of = [&](size_t c) { return std::min(J.ch, J.bytes - c * J.ch); };
    int k = 0;
    if (J.dir == 0) {
        bool used[2] = {false, false};
        for (;;) {
            if (J.fail.load(std::memory_order_relaxed)) return;
            const size_t c = J.next.fetch_add(1, std::memory_order_relaxed);
            if (c >= J.nch) return;
            if (used[k] && !wait(p, k)) break;
            memcpy(slot[k], J.host + c * J.ch, len_of(c));
            if (!issue(p, c, k)) break;
            used[k] = true;
            k ^= 1;
        }
    } else {
        size_t pend_c[2] = {0, 0};
        int pend_k[2] = {0, 0}, npend = 0;
        bool more = true;
        for (;;) {
            if (J.fail.load(std::memory_order_relaxed)) return;
            if (more) {
                const size_t c = J.next.fetch_add(1, std::memory_order_relaxed);
                if (c >= J.nch) {
                    more = false;
                } else {
                    if (!issue(p, c, k)) break;
                    pend_c[npend] = c, pend_k[npend] = k, npend++;
                    k ^= 1;
                }
            }
            if (npend == 2 || (!more && npend > 0)) {  // the older chunk: wait for it, copy it out
                if (!wait(p, pend_k[0])) break;
                memcpy(J.host + pend_c[0] * J.ch, slot[pend_k[0]], len_of(pend_c[0]));
                pend_c[0] = pend_c[1], pend_k[0] = pend_k[1], npend--;
            }
            if (!more && npend == 0) return;
        }
    }
    J.fail.store(1);
}

template <class Enter, class Issue, class Wait>
inline bool s3d_pipe2_run(Pipe2Job &J, Enter &&enter, Issue &&issue, Wait &&wait, unsigned first_worker = 0)
{
    HostTeam &team = HostTeam::get();
    const unsigned me = team.workers();
    team.run(first_worker, [&](unsigned t) { s3d_pipe2_participant(J, t, enter, issue, wait); },
             [&] { s3d_pipe2_participant(J, me, enter, issue, wait); });
    return J.fail.load() == 0;
}
}  // namespace
