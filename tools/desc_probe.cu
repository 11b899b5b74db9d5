// Hardware probe (run on a B200 through gpurun): how does tcgen05.mma address a SWIZZLE_128B K-major operand whose
// descriptor start address is shifted by whole 128-byte rows inside a 1024-byte swizzle atom?
//
// The halo-reuse 3x3 convolution keeps ONE activation patch per 64-channel chunk in shared memory (written by TMA or
// by transform warps in the TMA swizzle pattern, base 1024-byte aligned) and issues the nine taps as MMAs whose A
// descriptors start at patch + s * 128 bytes, s = dy * (W + 1) + dx.  This program loads a [160 x 64] fp16 matrix with
// TMA, runs D = A[s : s + 128, :] * B^T for s = 0..17 and every value of the descriptor's 3-bit "matrix base offset"
// field (bits 49-51), and reports which base-offset values reproduce the host result.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/desc_probe.bin tools/desc_probe.cu
#include "../ml-stable-diffusion_b200/csrc/common.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace b200sd;

namespace b200sd {
void set_error(const char*, ...) {}
}  // namespace b200sd

static constexpr int kRows = 160, kK = 64, kN = 16, kShifts = 18, kBo = 8;

struct __align__(64) ProbeParams {
    CUtensorMap tmA, tmB;
    float* out;  // [kShifts][kBo][128][kN]
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ ProbeParams p) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;                       // 160 rows * 128 B = 20480
    uint8_t* sb = smem + 24 * 1024;           // 16 rows * 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 28 * 1024);
    uint64_t* mbar = bar + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(mbar, 1);
        fence_barrier_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr, 32);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, kRows * 128 + kN * 128);
        tma_load_2d(sa, &p.tmA, bar, 0, 0, kEvictNormal);
        tma_load_2d(sb, &p.tmB, bar, 0, 0, kEvictNormal);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(128, kN, 0, 0);
    uint32_t phase = 0;
    for (int s = 0; s < kShifts; ++s) {
        for (int bo = 0; bo < kBo; ++bo) {
            if (threadIdx.x == 0) {
                const uint64_t adesc = make_smem_desc_sw128(smem_u32(sa) + s * 128, 1024, 0) | (static_cast<uint64_t>(bo) << 49);
                const uint64_t bdesc = make_smem_desc_sw128(smem_u32(sb), 1024, 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0 ? 1u : 0u);
                umma_commit(mbar);
            }
            mbar_wait(mbar, phase);
            phase ^= 1;
            tc_fence_after();
            uint32_t v[16];
            tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
            tmem_ld_wait();
            float* o = p.out + ((static_cast<size_t>(s) * kBo + bo) * 128 + warp * 32 + lane) * kN;
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
            tc_fence_before();
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 32);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn fn, CUtensorMap* m, void* base, int rows, int box_rows) {
    cuuint64_t gd[2] = {kK, static_cast<cuuint64_t>(rows)};
    cuuint64_t gs[1] = {kK * 2};
    cuuint32_t bx[2] = {kK, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t es[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(fp);
    std::vector<__half> ha(kRows * kK), hb(kN * kK);
    std::vector<float> fa(kRows * kK), fb(kN * kK);
    srand(7);
    for (int i = 0; i < kRows * kK; ++i) {
        fa[i] = static_cast<float>(rand() % 9 - 4);
        ha[i] = __float2half(fa[i]);
    }
    for (int i = 0; i < kN * kK; ++i) {
        fb[i] = static_cast<float>(rand() % 7 - 3);
        hb[i] = __float2half(fb[i]);
    }
    __half *da, *db;
    float* dout;
    const size_t out_n = static_cast<size_t>(kShifts) * kBo * 128 * kN;
    cudaMalloc(&da, ha.size() * 2);
    cudaMalloc(&db, hb.size() * 2);
    cudaMalloc(&dout, out_n * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, out_n * 4);
    ProbeParams p;
    if (make_map(fn, &p.tmA, da, kRows, kRows) || make_map(fn, &p.tmB, db, kN, kN)) {
        printf("tensor map encode failed\n");
        return 1;
    }
    p.out = dout;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    probe_kernel<<<1, 128, 40 * 1024>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("kernel failed: %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::vector<float> out(out_n);
    cudaMemcpy(out.data(), dout, out_n * 4, cudaMemcpyDeviceToHost);
    int all_zero_ok = 1, all_rel_ok = 1;
    for (int s = 0; s < kShifts; ++s) {
        printf("shift %2d rows: base_offset values that match:", s);
        int zero_ok = 0, rel_ok = 0;
        for (int bo = 0; bo < kBo; ++bo) {
            bool ok = true;
            for (int r = 0; r < 128 && ok; ++r)
                for (int n = 0; n < kN && ok; ++n) {
                    float ref = 0.f;
                    for (int k = 0; k < kK; ++k) ref += fa[(s + r) * kK + k] * fb[n * kK + k];
                    if (out[((static_cast<size_t>(s) * kBo + bo) * 128 + r) * kN + n] != ref) ok = false;
                }
            if (ok) {
                printf(" %d", bo);
                if (bo == 0) zero_ok = 1;
                if (bo == (s & 7)) rel_ok = 1;
            }
        }
        printf("\n");
        all_zero_ok &= zero_ok;
        all_rel_ok &= rel_ok;
    }
    printf("SUMMARY base_offset=0 always correct: %d ; base_offset=(shift&7) always correct: %d\n", all_zero_ok, all_rel_ok);
    return 0;
}
