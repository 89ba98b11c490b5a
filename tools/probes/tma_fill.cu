// Probe: which bit pattern does CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA write for fp32 / fp16 / bf16 maps?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_fill tools/probes/tma_fill.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__global__ void k(const __grid_constant__ CUtensorMap m, uint32_t* out) {
  __shared__ __align__(128) uint32_t tile[64];
  __shared__ __align__(8) uint64_t bar;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(tile);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(64u));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(d), "l"(&m), "r"(-4), "r"(-1), "r"(b) : "memory");
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra W;\n}" ::"r"(b) : "memory");
    for (int i = 0; i < 16; ++i) out[i] = tile[i];
  }
}
int main() {
  cuInit(0);
  void* buf; cudaMalloc(&buf, 4096); cudaMemset(buf, 0x11, 4096);
  uint32_t* out; cudaMallocManaged(&out, 64);
  CUtensorMapDataType types[3] = {CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16};
  const char* names[3] = {"f32", "f16", "bf16"};
  for (int t = 0; t < 3; ++t) {
    cuuint64_t esz = t == 0 ? 4 : 2;
    cuuint64_t dims[2] = {32, 8}, strides[1] = {32 * esz};
    cuuint32_t box[2] = {(cuuint32_t)(32 / esz), 2}, es[2] = {1, 1};  // 32-byte rows x 2 = 64 bytes
    CUtensorMap m;
    CUresult r = cuTensorMapEncodeTiled(&m, types[t], 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
    if (r != CUDA_SUCCESS) { printf("%s encode failed %d\n", names[t], (int)r); continue; }
    k<<<1, 32>>>(m, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s (%s):", names[t], cudaGetErrorString(e));
    for (int i = 0; i < 16; ++i) printf(" %08x", out[i]);
    printf("\n");
  }
  return 0;
}
