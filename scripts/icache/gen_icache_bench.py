"""Generator of an instruction-fetch microbenchmark (development tool, evidence for DESIGN.md section 4a).
Emits icache_bench.cu: straight-line FP64 code bodies of a given size, executed R times in a loop by W warps per CTA and
C CTAs per SM, either the SAME body for every warp or a DIFFERENT copy per warp.  Reports warp-instructions/clk/SM.
usage: python gen_icache_bench.py > icache_bench.cu; nvcc -O3 -arch=sm_100a -o icache_bench icache_bench.cu"""
import sys

SIZES_KB = [4, 8, 16, 24, 32, 48, 64, 96]
DIFF_KB = [2, 4, 8, 16]
NACC = 12

def body(n, salt):
    out = []
    for i in range(n):
        a = (i * 5 + salt) % NACC
        b = (i * 7 + 3 + salt) % NACC
        out.append(f"    a{a} = fma(a{a}, x, a{b});")
    return "\n".join(out)

print("#include <cstdio>\n#include <cstdlib>\n#include <cuda_runtime.h>\n")
decl = " ".join(f"double a{i} = t * {i + 1}.0e-3;" for i in range(NACC))
summ = " + ".join(f"a{i}" for i in range(NACC))
for kb in SIZES_KB:
    n = kb * 1024 // 16
    print(f"__global__ void __launch_bounds__(256, 1) same_{kb}(double *out, int R, double x)\n{{\n  extern __shared__ double sm[];\n"
          f"  const double t = threadIdx.x; {decl}\n#pragma unroll 1\n  for (int r = 0; r < R; r++) {{\n{body(n, 0)}\n  }}\n"
          f"  out[blockIdx.x * blockDim.x + threadIdx.x] = {summ};\n}}\n")
for kb in DIFF_KB:
    n = kb * 1024 // 16
    print(f"__global__ void __launch_bounds__(256, 1) diff_{kb}(double *out, int R, double x)\n{{\n  extern __shared__ double sm[];\n"
          f"  const double t = threadIdx.x; {decl}\n  const int w = threadIdx.x >> 5;\n#pragma unroll 1\n  for (int r = 0; r < R; r++) {{\n    switch (w) {{")
    for w in range(8):
        print(f"    case {w}:\n{body(n, w + 1)}\n      break;")
    print(f"    }}\n  }}\n  out[blockIdx.x * blockDim.x + threadIdx.x] = {summ};\n}}\n")

print(r"""
typedef void (*K)(double *, int, double);
static double run(K k, int ninstr, int W, int C, double *out, double clk_ghz, int sms)
{
  const int smem = (227 * 1024) / C - 1024 - 64;
  cudaFuncSetAttribute((const void *) k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long target = 6000000;                       // warp-instructions per warp per launch
  int R = (int) (target / ninstr); if (R < 2) R = 2;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<sms * C, 32 * W, smem>>>(out, 2, 1.0000001);   // warm-up
  cudaEventRecord(e0);
  k<<<sms * C, 32 * W, smem>>>(out, R, 1.0000001);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  if (cudaGetLastError() != cudaSuccess) { printf("launch failed\n"); return 0; }
  return (double) W * C * R * ninstr / (ms * 1e-3 * clk_ghz * 1e9);   // warp-instr / clk / SM
}
int main()
{
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz * 1e-6; const int sms = p.multiProcessorCount;
  double *out; cudaMalloc(&out, 1 << 24);
  printf("%s, %d SMs, %.3f GHz; FP64 pipe bound = 2 warp-DFMA/clk/SM\n", p.name, sms, ghz);
  const int WC[][2] = {{1, 1}, {2, 1}, {4, 1}, {8, 1}, {1, 2}, {1, 4}, {1, 8}, {4, 2}, {4, 4}};
""")
print("  struct { const char *name; K k; int n; } same[] = {" +
      ", ".join(f'{{"same_{kb}KB", same_{kb}, {kb * 1024 // 16}}}' for kb in SIZES_KB) + "};")
print("  struct { const char *name; K k; int n; } diff[] = {" +
      ", ".join(f'{{"diff_{kb}KB/warp", diff_{kb}, {kb * 1024 // 16}}}' for kb in DIFF_KB) + "};")
print(r"""
  printf("%-16s", "body \\ WxC");
  for (auto &wc : WC) printf(" %dx%d   ", wc[0], wc[1]);
  printf("\n");
  for (auto &s : same) {
    printf("%-16s", s.name);
    for (auto &wc : WC) printf(" %6.3f", run(s.k, s.n, wc[0], wc[1], out, ghz, sms));
    printf("\n");
  }
  const int WD[][2] = {{8, 1}, {8, 2}, {4, 1}, {4, 2}, {4, 4}, {2, 4}};
  printf("%-16s", "diff \\ WxC");
  for (auto &wc : WD) printf(" %dx%d   ", wc[0], wc[1]);
  printf("\n");
  for (auto &s : diff) {
    printf("%-16s", s.name);
    for (auto &wc : WD) printf(" %6.3f", run(s.k, s.n, wc[0], wc[1], out, ghz, sms));
    printf("\n");
  }
  return 0;
}
""")
