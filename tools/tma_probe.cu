// minimal TMA 4-D fp64 load probe: argv: boxk boxj nf dtype(0=f64,1=u64) offk offj
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, double* out, int n, int c0, int c1, int c2, unsigned bytes) {
  extern __shared__ __align__(128) unsigned char raw[];
  double* buf = reinterpret_cast<double*>(raw);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(raw + 32768);
  const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar);
  const unsigned dst = (unsigned)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(1u), "r"(mb) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(mb) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(&tm), "r"(mb), "r"(c0), "r"(c1), "r"(c2), "r"(0) : "memory");
  }
  asm volatile("{\n .reg .pred P1;\n W: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra D;\n bra W;\n D:\n}\n" ::"r"(mb), "r"(0u) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = buf[i];
}
int main(int argc, char** argv) {
  int bk = atoi(argv[1]), bj = atoi(argv[2]), nf = atoi(argv[3]), dt = atoi(argv[4]), ok = atoi(argv[5]), oj = atoi(argv[6]);
  const int nk = 257, nj = 257, ni = 5, pitch = 272;
  const long long plane = (long long)nj * pitch, field = ni * plane;
  std::vector<double> h(2 * field);
  for (long long i = 0; i < 2 * field; ++i) h[i] = (double)i;
  double* d; cudaMalloc(&d, sizeof(double) * 2 * field);
  cudaMemcpy(d, h.data(), sizeof(double) * 2 * field, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  CUtensorMap tm;
  cuuint64_t dims[4] = {(cuuint64_t)nk, (cuuint64_t)nj, (cuuint64_t)ni, (cuuint64_t)nf};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * 8, (cuuint64_t)plane * 8, (cuuint64_t)field * 8};
  cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)bj, 1u, (cuuint32_t)nf};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult rc = enc(&tm, dt ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, d, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)rc);
  const int n = bk * bj * nf;
  double* out; cudaMalloc(&out, sizeof(double) * n);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  k<<<1, 128, 40000>>>(tm, out, n, ok, oj, 1, (unsigned)(n * 8));
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) {
    std::vector<double> r(n); cudaMemcpy(r.data(), out, sizeof(double) * n, cudaMemcpyDeviceToHost);
    // expected element (f, jj, kk) = f*field + 1*plane + (oj+jj)*pitch + (ok+kk), 0 when out of bounds
    int bad = 0;
    for (int f = 0; f < nf; ++f) for (int jj = 0; jj < bj; ++jj) for (int kk = 0; kk < bk; ++kk) {
      int J = oj + jj, K = ok + kk;
      double ex = (J < 0 || J >= nj || K < 0 || K >= nk) ? 0.0 : (double)(f * field + plane + (long long)J * pitch + K);
      if (r[(f * bj + jj) * bk + kk] != ex) ++bad;
    }
    printf("mismatches %d of %d; first %.0f\n", bad, n, r[0]);
  }
  return 0;
}
