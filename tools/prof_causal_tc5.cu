// Phase timing of the tcgen05 causal attention kernel (clock64 stamps of the heaviest CTA).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -DFA_PROFILE -I. -o build_tmp/prof_causal_tc5 tools/prof_causal_tc5.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../psalm_b200/csrc/attn_tc5.cu"
namespace psalm { void set_error(const char*, ...) {} }
int main(int argc, char** argv) {
  const int T_ = argc > 1 ? atoi(argv[1]) : 1024, nh = 32, B = 1;
  size_t n = (size_t)B * T_ * 3 * nh * 64;
  std::vector<__nv_bfloat16> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = __float2bfloat16((float)rand() / RAND_MAX - 0.5f);
  __nv_bfloat16 *qkv, *out;
  cudaMalloc(&qkv, n * 2);
  cudaMalloc(&out, (size_t)B * T_ * nh * 64 * 2);
  cudaMemcpy(qkv, h.data(), n * 2, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int it = 0; it < 5; ++it) {
    long long z[16] = {0};
    cudaMemcpyToSymbol(fa_prof, z, sizeof(z));
    cudaEventRecord(e0);
    int rc = psalm::tc5_causal_attention(qkv, nullptr, out, B, T_, nh, 64, PSALM_BF16, 0);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long p[16];
    cudaMemcpyFromSymbol(p, fa_prof, sizeof(p));
    printf("rc=%d err=%s kernel %.1f us; heaviest CTA cycles by phase:", rc, cudaGetErrorString(e), ms * 1e3);
    const char* nm[] = {"wait_S", "ldS+waitPV", "barA", "max", "rescale+exp+P", "barB", "-", "-", "-"};
    long long tot = 0;
    for (int i = 0; i < 9; ++i) tot += p[i];
    for (int i = 0; i < 9; ++i) printf(" %s=%lld", nm[i], p[i]);
    printf(" total=%lld\n", tot);
  }
  return 0;
}
