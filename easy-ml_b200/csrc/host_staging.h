// host_staging.h -- moving pageable caller memory (a Rust Vec<T>, a numpy array) to and from the device at
// PCIe speed.  cudaMemcpyAsync on pageable memory is staged by the driver on one thread (~11 GB/s measured here,
// and it serialises with the compute); this stages through a ring of page-locked slots filled / drained by a small
// thread pool, so the copies stay asynchronous and overlap the GEMM like they do for page-locked buffers.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

#include "common.h"

namespace eml {

// true when cudaMemcpyAsync on this pointer is a true asynchronous DMA (page-locked or registered memory)
bool host_pointer_is_pinned(const void* p);

// dst[r*dpitch .. +width) = src[r*spitch .. +width) for r < rows, split over the copy pool's threads
void parallel_copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, int64_t rows);

class StagedCopier {
  public:
    explicit StagedCopier(DeviceCtx* ctx) : ctx_(ctx) {}
    ~StagedCopier() { finish(); }
    // 2-D copies with the semantics of cudaMemcpy2DAsync; the host side may be pageable
    int h2d(void* dst_dev, size_t dpitch, const void* src_host, size_t spitch, size_t width, int64_t rows,
            cudaStream_t stream);
    int d2h(void* dst_host, size_t dpitch, const void* src_dev, size_t spitch, size_t width, int64_t rows,
            cudaStream_t stream);
    // completes every pending device-to-host chunk (blocks until its data has landed in the caller's memory)
    int finish();

  private:
    struct Pending {
        int slot;
        void* dst;
        size_t dpitch, width;
        int64_t rows;
    };
    int acquire(int dir, int* slot);
    int complete(const Pending& p);
    int drain_ready();
    DeviceCtx* ctx_;
    int next_[2] = {0, 0};
    std::vector<Pending> pending_;  // device-to-host chunks whose data still sits in a slot
};

}  // namespace eml
