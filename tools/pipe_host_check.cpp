// pipe_host_check.cpp -- the chunk pipeline of sift3d_b200/csrc/host_pipe.h against a MOCK DMA
// engine on the CPU: a thread that executes the queued copies in order, each after a random
// delay, and marks the slot's "event" complete -- the behaviour pipe_transfer (engine.cu) gets
// from cudaMemcpyAsync + cudaEventRecord / cudaEventQuery on one stream.  Checks that every
// transfer delivers the exact bytes for sizes around the chunk / ring boundaries, in both
// directions, and that no slot is overwritten while the other side still needs it (the mock
// poisons a slot after an upload DMA has read it and checks a download slot is not refilled
// before its chunk was copied out).
//   g++ -O2 -std=c++17 -pthread -I sift3d_b200/csrc tools/pipe_host_check.cpp -o /tmp/pipe_host_check
#include "host_pipe.h"

#include <chrono>
#include <cstdio>
#include <deque>
#include <random>

namespace {
struct MockDma {
    struct Op {
        char *dst;
        const char *src;
        size_t len;
        std::atomic<int> *ev;
        char *poison;  // upload: the slot is dead after the DMA has read it
    };
    std::deque<Op> q;
    std::mutex mu;
    std::condition_variable cv;
    bool stop = false;
    unsigned max_delay_us;
    std::thread th;
    explicit MockDma(unsigned d) : max_delay_us(d), th([this] { loop(); }) {}
    ~MockDma()
    {
        {
            std::lock_guard<std::mutex> g(mu);
            stop = true;
        }
        cv.notify_all();
        th.join();
    }
    void push(Op op)
    {
        {
            std::lock_guard<std::mutex> g(mu);
            q.push_back(op);
        }
        cv.notify_one();
    }
    void loop()
    {
        std::mt19937 rng(7);
        for (;;) {
            Op op;
            {
                std::unique_lock<std::mutex> g(mu);
                cv.wait(g, [&] { return stop || !q.empty(); });
                if (q.empty()) return;
                op = q.front();
                q.pop_front();
            }
            if (max_delay_us) std::this_thread::sleep_for(std::chrono::microseconds(rng() % (max_delay_us + 1)));
            memcpy(op.dst, op.src, op.len);
            if (op.poison) memset(op.poison, 0xEE, op.len);
            op.ev->store(1, std::memory_order_release);
        }
    }
};

int run_case(int dir, size_t bytes, size_t ch, size_t ns, unsigned delay_us, unsigned seed, unsigned first_worker)
{
    std::vector<char> host(bytes), dev(bytes), slots(ch * ns), want(bytes);
    std::mt19937 rng(seed);
    for (size_t i = 0; i < bytes; i++) want[i] = (char)rng();
    if (dir == 0) host = want, std::fill(dev.begin(), dev.end(), 0x55);
    else dev = want, std::fill(host.begin(), host.end(), 0x55);
    std::vector<std::atomic<int>> ev(ns);
    for (auto &x : ev) x.store(1);
    PipeJob J;
    J.dir = dir;
    J.host = host.data();
    J.slots = slots.data();
    J.bytes = bytes, J.ch = ch, J.ns = ns;
    J.nch = (bytes + ch - 1) / ch;
    std::unique_ptr<std::atomic<unsigned char>[]> flags(new std::atomic<unsigned char>[J.nch]);
    for (size_t c = 0; c < J.nch; c++) flags[c].store(0);
    J.flag = flags.get();
    int bad = 0;
    {
        MockDma dma(delay_us);
        auto issue = [&](size_t c) {
            char *slot = J.slots + (c % ns) * ch;
            const size_t len = std::min(ch, bytes - c * ch);
            // the slot's previous DMA must be complete before its event is recorded again, and a
            // download slot must have been copied out before it is refilled
            if (!ev[c % ns].load(std::memory_order_acquire)) bad++;
            if (dir == 1 && c >= ns && !J.flag[c - ns].load(std::memory_order_acquire)) bad++;
            ev[c % ns].store(0, std::memory_order_release);
            if (dir == 0) dma.push({dev.data() + c * ch, slot, len, &ev[c % ns], slot});
            else dma.push({slot, dev.data() + c * ch, len, &ev[c % ns], nullptr});
            return true;
        };
        auto poll = [&](size_t c) { return ev[c % ns].load(std::memory_order_acquire) ? 1 : 0; };
        if (!s3d_pipe_run(J, issue, poll, first_worker)) bad++;
        // cudaStreamSynchronize: the queue drains before the transfer returns
        for (size_t s = 0; s < ns; s++)
            while (!ev[s].load(std::memory_order_acquire)) std::this_thread::yield();
    }
    const std::vector<char> &got = dir == 0 ? dev : host;
    if (memcmp(got.data(), want.data(), bytes)) bad++;
    if (J.ndone.load() != J.nch) bad++;
    return bad;
}

// The poller-free pipeline (Pipe2Job): one FIFO "stream" per participant, all served by the one
// mock DMA thread; an event per (participant, slot).
int run_case2(int dir, size_t bytes, size_t ch, unsigned delay_us, unsigned seed, unsigned first_worker)
{
    const unsigned P = HostTeam::get().workers() + 1;
    std::vector<char> host(bytes), dev(bytes), slots(ch * 2 * P), want(bytes);
    std::mt19937 rng(seed);
    for (size_t i = 0; i < bytes; i++) want[i] = (char)rng();
    if (dir == 0) host = want, std::fill(dev.begin(), dev.end(), 0x55);
    else dev = want, std::fill(host.begin(), host.end(), 0x55);
    std::vector<std::atomic<int>> ev(2 * P);
    for (auto &x : ev) x.store(1);
    Pipe2Job J;
    J.dir = dir;
    J.host = host.data();
    J.slots = slots.data();
    J.bytes = bytes, J.ch = ch;
    J.nch = (bytes + ch - 1) / ch;
    std::atomic<int> bad{0};
    {
        MockDma dma(delay_us);  // one queue: in order over all participants, a fortiori per participant
        auto enter = [&](unsigned) { return true; };
        auto issue = [&](unsigned p, size_t c, int k) {
            char *slot = J.slots + (size_t)(2 * p + k) * ch;
            const size_t len = std::min(ch, bytes - c * ch);
            if (!ev[2 * p + k].load(std::memory_order_acquire)) bad++;  // slot still in flight
            ev[2 * p + k].store(0, std::memory_order_release);
            if (dir == 0) dma.push({dev.data() + c * ch, slot, len, &ev[2 * p + k], slot});
            else dma.push({slot, dev.data() + c * ch, len, &ev[2 * p + k], nullptr});
            return true;
        };
        auto wait = [&](unsigned p, int k) {
            while (!ev[2 * p + k].load(std::memory_order_acquire)) std::this_thread::yield();
            return true;
        };
        if (!s3d_pipe2_run(J, enter, issue, wait, first_worker)) bad++;
        for (auto &x : ev)
            while (!x.load(std::memory_order_acquire)) std::this_thread::yield();
    }
    const std::vector<char> &got = dir == 0 ? dev : host;
    if (memcmp(got.data(), want.data(), bytes)) bad++;
    return bad.load();
}
}  // namespace

int main()
{
    int bad = 0, cases = 0;
    const size_t chs[] = {4096, 65536, 1 << 20};
    const size_t nss[] = {2, 3, 8, 16};
    for (int dir = 0; dir < 2; dir++)
        for (size_t ch : chs)
            for (size_t ns : nss) {
                const size_t sizes[] = {1, ch - 1, ch, ch + 1, ch * ns - 7, ch * ns, ch * ns + 1,
                                        ch * (2 * ns + 1) + 123, ch * 37 + 4095};
                for (size_t bytes : sizes) {
                    if (bytes > ((size_t)48 << 20)) continue;
                    const unsigned delay = ch <= 65536 ? (cases % 3 == 0 ? 0 : 30) : 0;
                    // several ranks on a host: the polling caller replaces worker 0 (engine.cu)
                    const unsigned first = HostTeam::get().workers() > 1 ? (unsigned)(cases & 1) : 0u;
                    const int b = run_case(dir, bytes, ch, ns, delay, 1000 + cases, first);
                    if (b) printf("FAIL dir %d bytes %zu ch %zu ns %zu\n", dir, bytes, ch, ns);
                    bad += b;
                    cases++;
                }
            }
    for (int dir = 0; dir < 2; dir++)
        for (size_t ch : chs) {
            const size_t sizes[] = {1, ch - 1, ch, ch + 1, 2 * ch, 2 * ch + 1, 7 * ch - 3, 40 * ch + 17, 200 * ch + 4095};
            for (size_t bytes : sizes) {
                if (bytes > ((size_t)48 << 20)) continue;
                const int b = run_case2(dir, bytes, ch, ch <= 65536 && (cases % 2) ? 20 : 0, 5000 + cases,
                                        HostTeam::get().workers() > 1 ? (unsigned)((cases >> 1) & 1) : 0u);
                if (b) printf("FAIL v2 dir %d bytes %zu ch %zu\n", dir, bytes, ch);
                bad += b;
                cases++;
            }
        }
    // the parallel memcpy of one chunk (HostTeam::copy), odd sizes
    for (size_t len : {(size_t)1, (size_t)4095, (size_t)4097, (size_t)1000003, (size_t)(8 << 20) + 5}) {
        std::vector<char> a(len), b(len, 0);
        for (size_t i = 0; i < len; i++) a[i] = (char)(i * 131 + 7);
        HostTeam::get().copy(b.data(), a.data(), len);
        if (memcmp(a.data(), b.data(), len)) bad++, printf("FAIL copy %zu\n", len);
        cases++;
    }
    printf("pipe_host_check: %d cases, workers %u, %s\n", cases, HostTeam::get().workers(), bad ? "FAILED" : "ok");
    return bad ? 1 : 0;
}
