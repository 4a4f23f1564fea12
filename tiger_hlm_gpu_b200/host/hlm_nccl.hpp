// hlm_nccl.hpp — the one inter-rank step of a routed run, for C++ hosts with one process per GPU: an NCCL
// all-gather of the boundary links' discharge per coupling interval, on the stream the library's kernels run on
// (nothing synchronises with the host between intervals).
//
// The reference's multi-process plumbing is MPI (main.cpp:258-309) and it couples no links at all
// (data/config.yaml:66-70 only sketches step buffers); this image has NCCL and no MPI, so ranks are plain
// processes with RANK / WORLD_SIZE / LOCAL_RANK in the environment (torchrun, or a shell loop) and the NCCL
// unique id travels through a file.  Used by hlm_run (routing.enabled with WORLD_SIZE > 1) and by the
// hlm_routed_nccl example.  Needs nccl.h, libnccl and the CUDA runtime; nothing else in the host tree does.
#pragma once

#include <cuda_runtime_api.h>
#include <nccl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <stdexcept>
#include <string>
#include <thread>

namespace hlm_b200 {

class NcclBoundaryExchange {
  public:
    /// world ranks, this one `rank` on CUDA device `device`; the id file lives in `id_dir` (a directory every rank
    /// sees).  max_send = the plan's per-rank segment length, halo_len = world * max_send.
    NcclBoundaryExchange(int world, int rank, int device, const std::string& id_dir, long long max_send, long long halo_len)
        : world_(world), rank_(rank), max_send_(max_send) {
        cuda(cudaSetDevice(device), "cudaSetDevice");
        cuda(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
        if (world_ <= 1) return;
        const std::string path = id_dir + "/" + id_file_name();
        const ncclUniqueId id = exchange_id(path);
        nccl(ncclCommInitRank(&comm_, world_, id, rank_), "ncclCommInitRank");
        // every rank holds the communicator once this tiny collective has gone through: the id file has done its job
        cuda(cudaMalloc(&d_flag_, sizeof(float)), "cudaMalloc");
        cuda(cudaMemsetAsync(d_flag_, 0, sizeof(float), stream_), "cudaMemset");
        nccl(ncclAllReduce(d_flag_, d_flag_, 1, ncclFloat, ncclSum, comm_, stream_), "ncclAllReduce");
        cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
        if (rank_ == 0) ::unlink(path.c_str());
        if (max_send_ > 0) {
            cuda(cudaMalloc(&d_send_, sizeof(double) * (size_t)max_send_), "cudaMalloc");
            cuda(cudaMalloc(&d_halo_, sizeof(double) * (size_t)halo_len), "cudaMalloc");
            cuda(cudaMemsetAsync(d_send_, 0, sizeof(double) * (size_t)max_send_, stream_), "cudaMemset");
            cuda(cudaMemsetAsync(d_halo_, 0, sizeof(double) * (size_t)halo_len, stream_), "cudaMemset");
        }
    }
    ~NcclBoundaryExchange() {
        if (comm_) ncclCommDestroy(comm_);
        if (d_send_) cudaFree(d_send_);
        if (d_halo_) cudaFree(d_halo_);
        if (d_flag_) cudaFree(d_flag_);
        if (stream_) cudaStreamDestroy(stream_);
    }
    NcclBoundaryExchange(const NcclBoundaryExchange&) = delete;
    NcclBoundaryExchange& operator=(const NcclBoundaryExchange&) = delete;

    cudaStream_t stream() const { return stream_; }
    double* d_send() const { return d_send_; }  // this rank's boundary discharge (the window kernel's epilogue writes it)
    double* d_halo() const { return d_halo_; }  // every rank's segment, rank order
    long long exchanges() const { return n_; }
    /// queue the all-gather of the padded per-rank segments behind the kernels already on the stream
    void all_gather() {
        if (world_ <= 1 || max_send_ <= 0) return;
        nccl(ncclAllGather(d_send_, d_halo_, (size_t)max_send_, ncclDouble, comm_, stream_), "ncclAllGather");
        ++n_;
    }
    /// a stream-ordered barrier (peer-memory exchange: the kernels have delivered the data themselves)
    void barrier() {
        if (world_ <= 1) return;
        nccl(ncclAllReduce(d_flag_, d_flag_, 1, ncclFloat, ncclSum, comm_, stream_), "ncclAllReduce");
        ++n_;
    }

  private:
    static void cuda(cudaError_t e, const char* what) {
        if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
    }
    static void nccl(ncclResult_t r, const char* what) {
        if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + ncclGetErrorString(r));
    }
    // One name per launch where the launcher says which launch this is (torchrun: TORCHELASTIC_RUN_ID / MASTER_PORT;
    // HLM_RUN_TOKEN for any other), so that two runs sharing a directory never see each other's id.
    static std::string id_file_name() {
        for (const char* key : {"HLM_RUN_TOKEN", "TORCHELASTIC_RUN_ID", "MASTER_PORT"}) {
            const char* v = std::getenv(key);
            if (v && *v) {
                std::string s(v);
                for (char& c : s)
                    if (!(std::isalnum((unsigned char)c) || c == '-' || c == '_')) c = '_';
                return "nccl_id_" + s + ".bin";
            }
        }
        return "nccl_id.bin";
    }
    // Rank 0 removes whatever an earlier run left, writes the id under a temporary name and renames it; the others
    // wait for the final name and take only a file written after they themselves started (less a few seconds for
    // ranks that start apart) — a stale id from a crashed run would hang ncclCommInitRank.  Rank 0 unlinks the file
    // once every rank holds the communicator (constructor).
    ncclUniqueId exchange_id(const std::string& path) const {
        ncclUniqueId id;
        if (rank_ == 0) {
            ::unlink(path.c_str());
            nccl(ncclGetUniqueId(&id), "ncclGetUniqueId");
            const std::string tmp = path + ".tmp";
            {
                std::ofstream f(tmp, std::ios::binary);
                f.write(reinterpret_cast<const char*>(&id), sizeof(id));
                if (!f) throw std::runtime_error("cannot write the NCCL id to " + tmp);
            }
            if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot publish the NCCL id at " + path);
            return id;
        }
        const std::time_t started = std::time(nullptr);
        for (int tries = 0; tries < 2400; ++tries) {
            struct stat st;
            if (::stat(path.c_str(), &st) == 0 && st.st_mtime >= started - 10) {
                std::ifstream f(path, std::ios::binary);
                if (f.read(reinterpret_cast<char*>(&id), sizeof(id))) return id;
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(50));
        }
        throw std::runtime_error("timed out waiting for the NCCL id at " + path);
    }

    int world_, rank_;
    long long max_send_, n_ = 0;
    cudaStream_t stream_ = nullptr;
    ncclComm_t comm_ = nullptr;
    double *d_send_ = nullptr, *d_halo_ = nullptr;
    float* d_flag_ = nullptr;
};

}  // namespace hlm_b200
