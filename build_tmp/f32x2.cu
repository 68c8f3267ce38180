#include <cstdint>
__global__ void k(const float2* a, float2* out, int n) {
  float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
  for (int i = threadIdx.x; i < n; i += 32) {
    float2 v = a[i];
    unsigned long long tp, q02, q13;
    asm("mov.b64 %0, {%1, %2};" : "=l"(tp) : "f"(v.x), "f"(v.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(q02) : "f"(acc0.x), "f"(acc0.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(q13) : "f"(acc1.x), "f"(acc1.y));
    asm("add.rn.f32x2 %0, %0, %1;" : "+l"(q02) : "l"(tp));
    asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(q13) : "l"(tp));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc0.x), "=f"(acc0.y) : "l"(q02));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(acc1.x), "=f"(acc1.y) : "l"(q13));
  }
  out[threadIdx.x] = make_float2(acc0.x + acc1.x, acc0.y + acc1.y);
}
