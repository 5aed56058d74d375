// firecode_b200 -- internal declarations shared by the embed pipelines (string / cyclical).
#pragma once

#include "fc_math.cuh"

namespace fc {

// pooled, stream-ordered device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaStream_t s = nullptr;
    cudaError_t alloc(size_t count, cudaStream_t stream) {
        release();
        s = stream;
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMallocAsync((void**)&p, count * sizeof(T), stream);
    }
    void release() {
        if (p) cudaFreeAsync(p, s);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { release(); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

// near-threshold decisions listed for the caller (north_star: "listed explicitly")
struct TieRecord {
    long long a;   // pose index (clash) or later pose of a pair (tfd / rmsd)
    long long b;   // -1 (clash) or earlier pose of the pair
    double value;  // min distance / torsion-difference sum / rmsd ...
    int kind;      // FC_TIE_*
    int decision;  // the decision the CUDA path took (1 = "below threshold")
};

// order-preserving compaction of the poses whose status has FC_STATUS_PASS set (CUB DeviceSelect)
int compact_pass(const uint8_t* status, long long n, long long base, long long* out_idx, int* out_count,
                 cudaStream_t s);

}  // namespace fc
