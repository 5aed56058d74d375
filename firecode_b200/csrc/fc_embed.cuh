// firecode_b200 -- internal declarations shared by the embed pipelines (string / cyclical).
#pragma once

#include <vector>

#include "fc_math.cuh"

// result object behind the opaque fc_result handle of the C-ABI (filled by the embed drivers)
struct fc_result {
    int64_t n_poses = 0, n_clash_pass = 0, n_rechecked = 0, n_kept = 0, n_atoms = 0, n_surv = 0, n_quads = 0;
    std::vector<uint8_t> status;         // per screened pose
    std::vector<int64_t> survivors;      // clash survivors (absolute pose index)
    std::vector<double> fingerprints;    // (n_surv, n_quads), stage-1 results only
    std::vector<int64_t> kept;           // absolute pose index, reference order
    std::vector<double> coords;          // (n_kept, n_atoms, 3)
    std::vector<int32_t> constrained;    // (n_kept, n_pairs, 2)
    int n_pairs = 0;
    std::vector<fc_tie> ties;
    int64_t ties_total = 0;
    // trimolecular cyclical embed: per group, the grid-search candidate _adjust_directions chose and
    // the cost gap to the runner-up (embeds.py:403-405)
    int64_t n_groups = 0;
    std::vector<int32_t> group_choice;
    std::vector<double> group_gap;
    // kept poses whose coordinates have NOT been materialised yet (trimolecular embed: 5.8 GB at BASELINE config C2):
    // per kept pose the rigid transform of every molecule and its conformer stay on the device, in one segment per
    // chunk of the screen; fc_result_kept_coords places the atoms slice by slice and streams them into the caller's
    // buffer.  Freed with the result.
    struct LazySegment {
        double* d_xf = nullptr;   // (count, n_mols, 12) R row-major, t
        int32_t* d_conf = nullptr;  // (count, n_mols)
        int64_t count = 0;
    };
    std::vector<LazySegment> lazy;
    double* lazy_coords[3] = {nullptr, nullptr, nullptr};  // device copies of the ensembles
    int32_t lazy_n_atoms[3] = {0, 0, 0};
    int32_t lazy_n_mols = 0;
    int lazy_device = -1;
    // string embed: coordinates of the kept poses stay on the device until fc_result_kept_coords copies them straight
    // into the caller's array (no intermediate host copy)
    double* d_kept_coords = nullptr;
    size_t d_kept_bytes = 0;
    ~fc_result();
};

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

fc_result* result_new();
// fc_host.cu: dst (host, pageable or pinned) <- src (device), through two pinned staging buffers drained by several
// host threads while the copy engine fills the other one
cudaError_t download_staged(void* dst, const void* src_dev, size_t bytes, cudaStream_t stream);

// fc_host.cu: pageable host rows -> device through two pinned staging buffers filled by several threads
cudaError_t upload_rows_staged(double* dst, const double* src, int64_t n, int n_atoms, const int32_t* sel, int n_sel,
                               cudaStream_t stream);

// order-preserving compaction of the poses whose status has FC_STATUS_PASS set (CUB DeviceSelect)
int compact_pass(const uint8_t* status, long long n, long long base, long long* out_idx, int* out_count,
                 cudaStream_t s);

}  // namespace fc
