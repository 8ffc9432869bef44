// nvcc 12.9 front-end bug found in round 2 (it cost one GPU call): inside a generic lambda,
//     if constexpr (K) { A } else for (...) { B }  S1;  S2;
// with the else branch an UNBRACED `for`, the first statement after it (S1) is dropped from the K = true
// instantiation (it is attached to the discarded else).  In film_conv3x3_tc*.cu S1 was `ra.advance(NA)`, the
// mbarrier ring position of the MMA warp -> the ring never advanced -> mbarrier timeouts.  Braces fix it.
// Check without a GPU:   nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -ptx <this file> -o - | grep st.global
// -> the store to [out+8] (ra.stage of the K = true instance) is the constant 0 although the loop ran n times.
#include <cstdio>
#include <type_traits>
struct R { int stage=0; __host__ __device__ void advance(int d){ if(++stage==d) stage=0; } };
__global__ void k(int n, int* out) {
  auto run = [&](auto tag) {
    constexpr bool K = decltype(tag)::value;
    R ra; int cnt = 0;
    for (int ab = 0; ab < n; ++ab) {
      if constexpr (K) {
        cnt += 10;
      } else
      for (int t = 0; t < 3; ++t) { cnt += 1; }
      ra.advance(100); cnt += 1000;
    }
    out[K ? 0 : 1] = cnt; out[K ? 2 : 3] = ra.stage;
  };
  run(std::true_type{}); run(std::false_type{});
}
int main(){ int* d; cudaMallocManaged(&d, 16); k<<<1,1>>>(5, d); cudaDeviceSynchronize(); printf("%d %d %d %d\n", d[0], d[1], d[2], d[3]); }
