// Host copy rate, pageable -> (ordinary) buffer, by thread count and store kind:
// gcc -O2 -pthread scripts/micro/host_copy.c -o /tmp/host_copy && /tmp/host_copy
// (is the staged pageable path of filter_host bound by memcpy's write-allocate traffic?)
#include <emmintrin.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct { char* dst; const char* src; size_t n; int nt; } job_t;
static void* run(void* p) {
  job_t* j = (job_t*)p;
  if (!j->nt) { memcpy(j->dst, j->src, j->n); return 0; }
  const __m128i* s = (const __m128i*)j->src;
  __m128i* d = (__m128i*)j->dst;
  size_t k = j->n / 16;
  for (size_t i = 0; i + 4 <= k; i += 4) {
    __m128i a = _mm_loadu_si128(s + i), b = _mm_loadu_si128(s + i + 1);
    __m128i c = _mm_loadu_si128(s + i + 2), e = _mm_loadu_si128(s + i + 3);
    _mm_stream_si128(d + i, a); _mm_stream_si128(d + i + 1, b);
    _mm_stream_si128(d + i + 2, c); _mm_stream_si128(d + i + 3, e);
  }
  _mm_sfence();
  return 0;
}
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(void) {
  const size_t total = (size_t)600 << 20, chunk = (size_t)28800000;
  char* src = aligned_alloc(4096, total); char* dst = aligned_alloc(4096, 3 * chunk + 4096);
  memset(src, 1, total); memset(dst, 2, 3 * chunk);
  for (int nt = 0; nt < 2; ++nt)
    for (int th = 1; th <= 16; th *= 2) {
      double best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        double t0 = now();
        size_t slot = 0;
        for (size_t off = 0; off + chunk <= total; off += chunk, slot = (slot + 1) % 3) {
          pthread_t tid[16]; job_t jb[16];
          size_t step = ((chunk / th) + 63) & ~(size_t)63;
          for (int i = 0; i < th; ++i) {
            size_t o = (size_t)i * step, n = o >= chunk ? 0 : (chunk - o < step ? chunk - o : step);
            jb[i] = (job_t){dst + slot * chunk + o, src + off + o, n, nt};
            pthread_create(&tid[i], 0, run, &jb[i]);
          }
          for (int i = 0; i < th; ++i) pthread_join(tid[i], 0);
        }
        double dt = now() - t0; if (dt < best) best = dt;
      }
      printf("%s stores, %2d threads: %.1f GB/s\n", nt ? "non-temporal" : "memcpy      ", th,
             (total / chunk) * chunk / best / 1e9);
    }
  return 0;
}
