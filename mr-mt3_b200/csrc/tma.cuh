// Tensor-map (TMA) plumbing shared by the tcgen05 GEMM, the decode-step projections and the decode
// attention: a host-side cache of CUtensorMap descriptors and the 2-D bulk-tensor load.
#pragma once
#include <cuda.h>

#include <map>
#include <tuple>

#include "common.cuh"

namespace mrmt3 {

constexpr int kTmaBoxCols = 64;  // every box is 64 bf16 (128 B) wide: one 128-byte-swizzle row

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            smem_dst),
        "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
}

// ---- host side: tensor maps -------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

class TmaCache {
public:
    // 2-D bf16 row-major tensor (rows, cols) with row pitch ld elements; box = box_rows x 64 cols
    Status get(const void* ptr, long rows, int cols, int ld, int box_rows, const CUtensorMap** out) {
        Key key{ptr, rows, cols, ld, box_rows};
        auto it = maps_.find(key);
        if (it == maps_.end()) {
            if (!encode_) {
                void* fn = nullptr;
                cudaDriverEntryPointQueryResult qres;
                MRMT3_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
                if (!fn || qres != cudaDriverEntryPointSuccess) return Error(2, "cuTensorMapEncodeTiled not available");
                encode_ = reinterpret_cast<PFN_tmapEncodeTiled>(fn);
            }
            CUtensorMap m;
            cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
            cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
            cuuint32_t box[2] = {(cuuint32_t)kTmaBoxCols, (cuuint32_t)box_rows};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = encode_(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
                                 estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return Error(2, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
            // bound the cache without invalidating pointers handed out recently: the full map becomes
            // the previous generation (its nodes keep their addresses) and is only destroyed at the
            // NEXT rotation, thousands of lookups later
            if (maps_.size() > 4096) {
                old_ = std::move(maps_);
                maps_.clear();
            }
            it = maps_.emplace(key, m).first;
        }
        *out = &it->second;
        return OkStatus();
    }
    void clear() {
        maps_.clear();
        old_.clear();
    }

private:
    typedef std::tuple<const void*, long, int, int, int> Key;
    std::map<Key, CUtensorMap> maps_, old_;
    PFN_tmapEncodeTiled encode_ = nullptr;
};


}  // namespace mrmt3
