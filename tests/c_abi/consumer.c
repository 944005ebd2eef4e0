/* A plain-C consumer of include/gdl_b200.h — what a non-Python host (or the reference's own C extension, if it had one) would write.
 * Test infrastructure: tests/test_c_abi_consumer_cpu.py compiles it with gcc -std=c99 and links it against the HOST-compiled library
 * (tests/hostemu), so the header is checked to be C-clean and the entry points to work with plain malloc'ed buffers: no torch, no
 * C++ types.  On a GPU box the same program, linked against libgdlb200.so and fed device pointers, does the same on the device.
 *
 * Path exercised: patch normalisation (utils/tensors.py:10-35) -> 5-class logits -> argmax (segmentation_segformer.py:268-271) ->
 * per-sample confusion counts (MeanIoU), plus the error convention (status code + gdl_last_error). */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gdl_b200.h"

static float bf16_to_float(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

#define CHECK(cond)                                                      \
  do {                                                                   \
    if (!(cond)) {                                                       \
      fprintf(stderr, "consumer.c:%d: %s failed (%s)\n", __LINE__, #cond, gdl_last_error()); \
      return 1;                                                          \
    }                                                                    \
  } while (0)

int main(void) {
  enum { N = 2, H = 4, W = 6, C = 3, LD = 8, K = 5 };
  CHECK(gdl_version() == GDL_B200_VERSION);

  /* 1. uint8 NHWC tile -> ((x / 255) - mean) / std as bf16 NHWC with the channels padded to 8 */
  uint8_t* raw = (uint8_t*)malloc(N * H * W * C);
  for (int i = 0; i < N * H * W * C; ++i) raw[i] = (uint8_t)((i * 37 + 11) & 255);
  const float mean[C] = {0.5f, 0.4f, 0.3f}, stdv[C] = {0.2f, 0.25f, 0.3f};
  uint16_t* y = NULL;
  CHECK(posix_memalign((void**)&y, 16, N * H * W * LD * sizeof(uint16_t)) == 0);
  CHECK(gdl_normalize_to_nhwc(raw, 0, y, GDL_BF16, N, H, W, C, LD, mean, stdv, 255.0f, NULL) == GDL_OK);
  for (int p = 0; p < N * H * W; ++p)
    for (int c = 0; c < LD; ++c) {
      const float got = bf16_to_float(y[p * LD + c]);
      const float want = c < C ? ((float)raw[p * C + c] / 255.0f - mean[c]) / stdv[c] : 0.0f;
      CHECK(fabsf(got - want) <= fabsf(want) * (1.0f / 128.0f) + 1e-6f); /* one bf16 rounding */
    }

  /* 2. logits -> class map, and the confusion counts against a target */
  float* logits = (float*)malloc(N * H * W * K * sizeof(float));
  long long* target = (long long*)malloc(N * H * W * sizeof(long long));
  for (int p = 0; p < N * H * W; ++p) {
    for (int k = 0; k < K; ++k) logits[p * K + k] = (float)((p * 7 + k * 13) % 17) - 8.0f;
    target[p] = (p * 3) % K;
  }
  long long* cls = (long long*)malloc(N * H * W * sizeof(long long));
  CHECK(gdl_argmax_classes(logits, K, (long long)N * H * W, K, 0.5f, cls, NULL) == GDL_OK);
  long long* conf = (long long*)calloc(N * K * K, sizeof(long long));
  long long* cls2 = (long long*)malloc(N * H * W * sizeof(long long));
  CHECK(gdl_argmax_confusion(logits, K, N, (long long)H * W, K, 0.5f, target, 0, -100, 0, cls2, conf, NULL) == GDL_OK);
  long long total = 0;
  for (int p = 0; p < N * H * W; ++p) {
    int best = 0;
    for (int k = 1; k < K; ++k)
      if (logits[p * K + k] > logits[p * K + best]) best = k; /* first maximum wins, as torch.argmax */
    CHECK(cls[p] == best && cls2[p] == best);
  }
  for (int n = 0; n < N; ++n)
    for (int t = 0; t < K; ++t)
      for (int k = 0; k < K; ++k) {
        long long want = 0;
        for (int p = n * H * W; p < (n + 1) * H * W; ++p) want += (target[p] == t && cls[p] == k);
        CHECK(conf[(n * K + t) * K + k] == want);
        total += want;
      }
  CHECK(total == N * H * W);

  /* 3. error convention: status code, message through gdl_last_error, nothing written */
  CHECK(gdl_normalize_to_nhwc(raw, 0, y, GDL_BF16, N, H, W, C, 2 /* pixel stride < channels */, mean, stdv, 255.0f, NULL) == GDL_ERR_INVALID);
  CHECK(strlen(gdl_last_error()) > 0);
  CHECK(gdl_normalize_to_nhwc(raw, 0, y, GDL_BF16, N, H, W, C, LD, mean, NULL /* mean without std */, 255.0f, NULL) == GDL_ERR_INVALID);
  CHECK(gdl_argmax_classes(NULL, K, 10, K, 0.5f, cls, NULL) == GDL_ERR_INVALID);

  printf("consumer.c: ok (%s)\n", "normalise -> argmax -> confusion counts through the C ABI");
  free(raw); free(y); free(logits); free(target); free(cls); free(cls2); free(conf);
  return 0;
}
