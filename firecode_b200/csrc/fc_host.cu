// firecode_b200 -- host-side data movement helpers (no kernels).
//
// upload_rows_staged: pageable numpy arrays reach the device at ~11 GB/s through cudaMemcpy (the driver stages
// them single-threaded); here a few host threads gather the wanted atoms of each structure into two pinned
// buffers while the copy engine sends the previous buffer, which moves the 576 MB of BASELINE config C4
// (only the heavy atoms of it) in a third of the time.
// fc_take_rows: the `structures[mask]` copy every pruning entry point returns
// (/root/reference/firecode/embedder.py:1400-1408 consumes it), done by several threads.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "fc_embed.cuh"

namespace fc {

// Worker threads for the host-side copies: one per physical core, at most 16 (FC_HOST_THREADS overrides).  The number of
// hardware threads per core comes from sysfs; boxes that expose one thread per core (the B200 pool's VMs) use them all.
static int host_threads() {
    static const int n = []() {
        if (const char* v = getenv("FC_HOST_THREADS"))
            if (atoi(v) > 0) return std::min(64, atoi(v));
        unsigned hc = std::thread::hardware_concurrency();
        if (!hc) return 1;
        unsigned per_core = 1;
        if (FILE* f = fopen("/sys/devices/system/cpu/cpu0/topology/thread_siblings_list", "r")) {
            char buf[64] = "";
            if (fgets(buf, sizeof buf, f)) {
                per_core = 1;
                for (const char* c = buf; *c; ++c)
                    if (*c == ',' || *c == '-') per_core = 2;
            }
            fclose(f);
        }
        return (int)std::max(1u, std::min(16u, hc / per_core));
    }();
    return n;
}

// memcpy with non-temporal stores for the big one-way copies of this file (staging buffers the copy engine reads next,
// result arrays the caller reads much later): the destination lines are not read into the cache first, which saves a
// third of the memory traffic of a plain copy.  Falls back to memcpy for unaligned or short pieces.
static inline void copy_stream(void* dst, const void* src, size_t bytes) {
#if defined(__SSE2__)
    if (bytes >= 256 && ((reinterpret_cast<uintptr_t>(dst) | bytes) & 15u) == 0) {
        __m128i* d = static_cast<__m128i*>(dst);
        const __m128i* s = static_cast<const __m128i*>(src);
        const size_t n16 = bytes / 16;
        size_t i = 0;
        for (; i + 4 <= n16; i += 4) {
            const __m128i a = _mm_loadu_si128(s + i), b = _mm_loadu_si128(s + i + 1);
            const __m128i c = _mm_loadu_si128(s + i + 2), e = _mm_loadu_si128(s + i + 3);
            _mm_stream_si128(d + i, a);
            _mm_stream_si128(d + i + 1, b);
            _mm_stream_si128(d + i + 2, c);
            _mm_stream_si128(d + i + 3, e);
        }
        for (; i < n16; ++i) _mm_stream_si128(d + i, _mm_loadu_si128(s + i));
        return;
    }
#endif
    memcpy(dst, src, bytes);
}
static inline void copy_stream_fence() {
#if defined(__SSE2__)
    _mm_sfence();  // the streamed lines are globally visible before anyone is told the copy is done
#endif
}

template <class F>
static void parallel_ranges(int64_t n, int64_t min_per_thread, F fn) {
    int nt = (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, n / std::max<int64_t>(1, min_per_thread)));
    if (nt <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t) {
        const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        th.emplace_back([=]() { fn(lo, hi); });
    }
    for (auto& x : th) x.join();
}

struct PinnedPair {
    void* p[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    size_t bytes = 0;
    int device = -1;  // the events belong to this device
};
static const size_t kStageBytes = (size_t)48 << 20;

// dst (device, n x n_sel x 3 doubles) <- src (host, n x n_atoms x 3 doubles) restricted to the atoms `sel`
// (null = all atoms).  Asynchronous on `stream` except for the staging itself.
//
// A ring of kUpSlots pinned slots: worker threads (alive for the whole call) fill slot c % kUpSlots with the rows of
// piece c, each its share of the rows, while the copy engine sends the pieces before it; the calling thread hands out
// slots whose previous copy has finished and submits filled pieces in order.
static const int kUpSlots = 4;
static const size_t kUpSlotBytes = (size_t)16 << 20;
struct UploadRing {
    char* base = nullptr;
    cudaEvent_t ev[kUpSlots] = {};
    int device = -1;
};

cudaError_t upload_rows_staged(double* dst, const double* src, int64_t n, int n_atoms, const int32_t* sel, int n_sel,
                               cudaStream_t stream) {
    static thread_local UploadRing st;
    if (n <= 0) return cudaSuccess;
    int dev = 0;
    cudaError_t e0 = cudaGetDevice(&dev);
    if (e0 != cudaSuccess) return e0;
    if (!st.base) {
        cudaError_t e = cudaHostAlloc((void**)&st.base, kUpSlots * kUpSlotBytes, cudaHostAllocPortable);
        if (e != cudaSuccess) return e;
    }
    if (st.device != dev) {  // first use, or the thread moved to another device: events are per device
        for (int b = 0; b < kUpSlots; ++b) {
            if (st.ev[b]) cudaEventDestroy(st.ev[b]);
            st.ev[b] = nullptr;
            cudaError_t e = cudaEventCreateWithFlags(&st.ev[b], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        st.device = dev;
    }
    const size_t row_out = (size_t)n_sel * 24, row_in = (size_t)n_atoms * 24;
    // the selection as runs of consecutive atoms {first atom, atoms}: heavy atoms usually sit together, so a row is a
    // handful of memcpy calls instead of one per atom
    std::vector<std::pair<int, int>> runs;
    for (int k = 0; sel && k < n_sel; ++k) {
        if (!runs.empty() && runs.back().first + runs.back().second == sel[k]) ++runs.back().second;
        else runs.emplace_back(sel[k], 1);
    }
    const int64_t rows_per_piece = std::max<int64_t>(1, (int64_t)(kUpSlotBytes / row_out));
    if ((size_t)rows_per_piece * row_out > kUpSlotBytes) return cudaErrorInvalidValue;  // one row larger than a slot
    const int64_t n_pieces = (n + rows_per_piece - 1) / rows_per_piece;
    const int nt = (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, std::min(n, rows_per_piece) / 512));

    char* const ring = st.base;  // `st` is thread-local: the workers must not name it
    std::mutex mu;
    std::condition_variable cv;
    int64_t released = 0;                              // pieces [0, released) may be filled
    std::vector<int> filled((size_t)n_pieces, 0);      // workers done with piece c
    auto fill = [&](int t) {
        for (int64_t c = 0; c < n_pieces; ++c) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return released > c || released < 0; });
                if (released < 0) return;  // the call failed
            }
            const int64_t r0 = c * rows_per_piece, rows = std::min(rows_per_piece, n - r0);
            const int64_t lo = rows * t / nt, hi = rows * (t + 1) / nt;
            char* stage = ring + (size_t)(c % kUpSlots) * kUpSlotBytes;
            for (int64_t r = lo; r < hi; ++r) {
                const char* in = (const char*)src + (size_t)(r0 + r) * row_in;
                char* out = stage + (size_t)r * row_out;
                if (!sel) {
                    copy_stream(out, in, row_out);
                } else {
                    for (const auto& run : runs) {
                        copy_stream(out, in + (size_t)run.first * 24, (size_t)run.second * 24);
                        out += (size_t)run.second * 24;
                    }
                }
            }
            copy_stream_fence();
            {
                std::lock_guard<std::mutex> lk(mu);
                ++filled[(size_t)c];
            }
            cv.notify_all();
        }
    };
    const bool trace = getenv("FC_PRUNE_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    double t_wait_fill = 0, t_wait_slot = 0;
    std::vector<std::thread> workers;
    workers.reserve((size_t)nt);
    for (int t = 0; t < nt; ++t) workers.emplace_back(fill, t);
    cudaError_t err = cudaSuccess;
    int64_t next_release = 0;
    // slot of piece r is free once the copy of piece r - kUpSlots has left the host (an event never recorded is complete)
    auto release_upto = [&](int64_t limit) {
        while (next_release < n_pieces && next_release < limit && err == cudaSuccess) {
            const double tw = now();
            err = cudaEventSynchronize(st.ev[next_release % kUpSlots]);
            t_wait_slot += now() - tw;
            ++next_release;
            {
                std::lock_guard<std::mutex> lk(mu);
                released = err == cudaSuccess ? next_release : -1;
            }
            cv.notify_all();
        }
    };
    release_upto(kUpSlots - 1);
    for (int64_t c = 0; c < n_pieces && err == cudaSuccess; ++c) {
        {
            const double tw = now();
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return filled[(size_t)c] == nt; });
            t_wait_fill += now() - tw;
        }
        const int64_t r0 = c * rows_per_piece, rows = std::min(rows_per_piece, n - r0);
        err = cudaMemcpyAsync((char*)dst + (size_t)r0 * row_out, st.base + (size_t)(c % kUpSlots) * kUpSlotBytes,
                              (size_t)rows * row_out, cudaMemcpyHostToDevice, stream);
        if (err == cudaSuccess) err = cudaEventRecord(st.ev[c % kUpSlots], stream);
        // piece c is in flight: waiting here for piece c - 2 never leaves the copy engine idle
        if (err == cudaSuccess) release_upto(c + kUpSlots - 1);
    }
    if (err != cudaSuccess) {
        std::lock_guard<std::mutex> lk(mu);
        released = -1;
    }
    cv.notify_all();
    for (auto& w : workers) w.join();
    if (trace)
        fprintf(stderr, "  upload_rows_staged: %lld rows, %.1f MB in %lld pieces, %d threads: %.2f ms on the host (waited %.2f ms for the "
                "threads to fill, %.2f ms for the copy engine to free slots)\n", (long long)n, 1e-6 * (double)n * (double)row_out,
                (long long)n_pieces, nt, now() - t_begin, t_wait_fill, t_wait_slot);
    return err;
}

}  // namespace fc

using namespace fc;

extern "C" int fc_take_rows(const void* src, int64_t row_bytes, const uint8_t* mask, int64_t n, void* dst, int64_t n_dst) {
    FC_REQUIRE(n >= 0 && row_bytes >= 0 && n_dst >= 0, "fc_take_rows: bad sizes");
    if (n == 0 || row_bytes == 0) return FC_OK;
    FC_REQUIRE(src && mask && (dst || n_dst == 0), "fc_take_rows: null pointer");
    // destination row of every kept source row (prefix sum), then a parallel copy
    std::vector<int64_t> kept;
    kept.reserve((size_t)n_dst);
    for (int64_t i = 0; i < n; ++i)
        if (mask[i]) kept.push_back(i);
    FC_REQUIRE((int64_t)kept.size() == n_dst, "fc_take_rows: mask selects %lld rows, destination holds %lld",
               (long long)kept.size(), (long long)n_dst);
    parallel_ranges(n_dst, 256, [&](int64_t lo, int64_t hi) {
        for (int64_t j = lo; j < hi; ++j)
            copy_stream((char*)dst + (size_t)j * (size_t)row_bytes, (const char*)src + (size_t)kept[(size_t)j] * (size_t)row_bytes,
                        (size_t)row_bytes);
        copy_stream_fence();
    });
    return FC_OK;
}

// ---- xyz text of a batch of structures, byte-identical to the reference's write_xyz ---------------
// (/root/reference/firecode/utils.py:105-116: "<n>\n<title>\n" then one line "%s     % .6f % .6f % .6f\n" per atom)
extern "C" int fc_xyz_format(const char* symbols, int32_t sym_stride, const double* coords, int64_t n, int32_t n_atoms,
                             const char* titles, char* out, int64_t out_cap, int64_t* out_len) {
    FC_REQUIRE(n >= 0 && n_atoms >= 0 && sym_stride > 0 && out_len, "fc_xyz_format: bad arguments");
    *out_len = 0;
    if (n == 0) return FC_OK;
    FC_REQUIRE((symbols && coords) || n_atoms == 0, "fc_xyz_format: null pointer");
    // titles: n NUL-terminated strings back to back (null -> "temp", the reference's default)
    std::vector<const char*> title((size_t)n, "temp");
    if (titles) {
        const char* t = titles;
        for (int64_t s = 0; s < n; ++s) {
            title[(size_t)s] = t;
            t += strlen(t) + 1;
        }
    }
    const int nt = (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, n * (int64_t)std::max(n_atoms, 1) / 4096));
    std::vector<std::string> part((size_t)nt);
    auto work = [&](int t) {
        const int64_t lo = n * t / nt, hi = n * (t + 1) / nt;
        std::string& buf = part[(size_t)t];
        buf.reserve((size_t)(hi - lo) * ((size_t)n_atoms * 48 + 32));
        char line[256];
        for (int64_t s = lo; s < hi; ++s) {
            buf += std::to_string(n_atoms);
            buf += '\n';
            buf += title[(size_t)s];
            buf += '\n';
            const double* x = coords + (size_t)s * n_atoms * 3;
            for (int k = 0; k < n_atoms; ++k) {
                const char* sym = symbols + (size_t)k * sym_stride;
                const int sl = (int)strnlen(sym, (size_t)sym_stride);
                const int len = snprintf(line, sizeof line, "%.*s     % .6f % .6f % .6f\n", sl, sym, x[3 * k], x[3 * k + 1], x[3 * k + 2]);
                buf.append(line, (size_t)std::min<int>(len, (int)sizeof line - 1));
            }
        }
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    int64_t total = 0;
    for (const std::string& b : part) total += (int64_t)b.size();
    *out_len = total;   // the size needed, also when the buffer is too small
    if (!out || out_cap < total) {
        set_error("fc_xyz_format: output needs %lld bytes, buffer holds %lld", (long long)total, (long long)out_cap);
        return FC_ERR_INVALID;
    }
    char* dst = out;
    for (const std::string& b : part) {
        memcpy(dst, b.data(), b.size());
        dst += b.size();
    }
    return FC_OK;
}

namespace fc {

// dst (host) <- src_dev (device): pieces of kStageBytes travel into two pinned staging buffers; while the copy engine
// fills one, several host threads copy the other into dst (for a fresh numpy array that is also where its pages are
// first touched).  Synchronous: returns when dst holds everything.
cudaError_t download_staged(void* dst, const void* src_dev, size_t bytes, cudaStream_t stream) {
    static thread_local PinnedPair st;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!st.p[0]) {
        for (int b = 0; b < 2; ++b) {
            e = cudaHostAlloc(&st.p[b], kStageBytes, cudaHostAllocPortable);
            if (e != cudaSuccess) return e;
        }
        st.bytes = kStageBytes;
    }
    if (st.device != dev) {
        for (int b = 0; b < 2; ++b) {
            if (st.ev[b]) cudaEventDestroy(st.ev[b]);
            e = cudaEventCreateWithFlags(&st.ev[b], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        st.device = dev;
    }
    const char* src = static_cast<const char*>(src_dev);
    char* out = static_cast<char*>(dst);
    size_t pend_off[2] = {0, 0}, pend_n[2] = {0, 0};
    int b = 0;
    for (size_t off = 0; off < bytes || pend_n[0] || pend_n[1]; b ^= 1) {
        if (pend_n[b]) {  // the piece that went into this staging buffer two steps ago
            e = cudaEventSynchronize(st.ev[b]);
            if (e != cudaSuccess) return e;
            const char* from = static_cast<const char*>(st.p[b]);
            char* to = out + pend_off[b];
            parallel_ranges((int64_t)pend_n[b], (int64_t)1 << 20, [=](int64_t lo, int64_t hi) { memcpy(to + lo, from + lo, (size_t)(hi - lo)); });
            pend_n[b] = 0;
        }
        if (off < bytes) {
            const size_t n = std::min(kStageBytes, bytes - off);
            e = cudaMemcpyAsync(st.p[b], src + off, n, cudaMemcpyDeviceToHost, stream);
            if (e != cudaSuccess) return e;
            e = cudaEventRecord(st.ev[b], stream);
            if (e != cudaSuccess) return e;
            pend_off[b] = off;
            pend_n[b] = n;
            off += n;
        }
    }
    return cudaSuccess;
}

}  // namespace fc

// ------------------------------------------------------------------------------------------------
// page-locked host memory on the GPU's NUMA node
// ------------------------------------------------------------------------------------------------
#include <sys/syscall.h>
#include <unistd.h>

namespace fc {

// NUMA node of the current device from sysfs (-1: unknown / single node)
static int device_numa_node() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), dev) != cudaSuccess) return -1;
    for (char* c = bus; *c; ++c)
        if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

}  // namespace fc

extern "C" void* fc_host_alloc(int64_t bytes, int32_t* node_out) {
    if (node_out) *node_out = -1;
    if (bytes <= 0) return nullptr;
    const int node = fc::device_numa_node();
    bool bound = false;
#ifdef SYS_set_mempolicy
    if (node >= 0 && node < 64) {  // MPOL_PREFERRED = 1: pages of this thread's next allocations come from `node`
        unsigned long mask = 1ul << node;
        bound = syscall(SYS_set_mempolicy, 1, &mask, 65ul) == 0;
    }
#endif
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocPortable);
    if (e == cudaSuccess && p) memset(p, 0, (size_t)bytes);  // first touch while the policy is in force
#ifdef SYS_set_mempolicy
    if (bound) syscall(SYS_set_mempolicy, 0, nullptr, 0ul);  // MPOL_DEFAULT
#endif
    if (e != cudaSuccess) {
        fc::cuda_fail(e, "cudaHostAlloc", __FILE__, __LINE__);
        return nullptr;
    }
    if (node_out) *node_out = bound ? node : -1;
    return p;
}

extern "C" void fc_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
