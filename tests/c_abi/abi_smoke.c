/* Plain-C consumer of include/crnn_ctc.h: what a maintainer binding libcrnnctc.so from C (or cgo / JNI / N-API glue) compiles.
 * Built and run by tests/test_api_cpu.py::test_c_abi_is_usable_from_plain_c with gcc -std=c99 -- no CUDA headers, no C++.
 * Exercises the entry points that need no GPU: version / status strings, the host beam-search decoder on TensorFlow's
 * testCTCDecoderBeamSearch vector (tests/golden/third_party_kats.py), the host copy pool, and the loud failure of
 * crnn_model_create on a box without a CUDA device.  Prints one "ok ..." line per check; exit code 0 only if all hold. */
#include <math.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "crnn_ctc.h"

static int failures = 0;
#define CHECK(cond, what)                                                        \
  do {                                                                           \
    if (cond) printf("ok   %s\n", what);                                          \
    else { printf("FAIL %s\n", what); ++failures; }                              \
  } while (0)

int main(int argc, char** argv) {
  const int expect_gpu = (argc > 1 && strcmp(argv[1], "--gpu") == 0);
  /* struct layout as this C compiler sees it; the Python test compares it with the ctypes mirror (lstm_ctc_ocr_b200/_lib.py) */
  printf("layout crnn_config %u %u %u %u %u %u %u\n", (unsigned)sizeof(crnn_config), (unsigned)offsetof(crnn_config, img_height),
         (unsigned)offsetof(crnn_config, nclasses), (unsigned)offsetof(crnn_config, num_hid), (unsigned)offsetof(crnn_config, bn_eps),
         (unsigned)offsetof(crnn_config, weight_decay), (unsigned)offsetof(crnn_config, compute_dtype));
  CHECK(crnn_version() > 0, "crnn_version");
  CHECK(strlen(crnn_status_string(CRNN_OK)) > 0 && strcmp(crnn_status_string(CRNN_OK), crnn_status_string(CRNN_INVALID_VALUE)) != 0,
        "crnn_status_string");

  /* beam search: 6 classes, blank 5, 5 of 8 frames valid, log p + 2.0; top path [1,0] at width 2, [0,1,0] at width 100 */
  {
    static const double p[6][6] = {
        {0.30999, 0.309938, 0.0679938, 0.0673362, 0.0708352, 0.173908},  {0.215136, 0.439699, 0.0370931, 0.0393967, 0.0381581, 0.230517},
        {0.199959, 0.489485, 0.0233221, 0.0251417, 0.0233289, 0.238763}, {0.279611, 0.452966, 0.0204795, 0.0209126, 0.0194803, 0.20655},
        {0.51286, 0.288951, 0.0243026, 0.0220788, 0.0219297, 0.129878},  {0.155251, 0.164444, 0.173517, 0.176138, 0.169979, 0.160671}};
    float x[8 * 6];
    int len = 5, out[8], out_len = -1, t, c;
    float nlp = 0.f;
    memset(x, 0, sizeof x);
    for (t = 0; t < 6; ++t) for (c = 0; c < 6; ++c) x[t * 6 + c] = (float)(log(p[t][c]) + 2.0);
    CHECK(crnn_ctc_beam_search(x, &len, 8, 1, 6, 2, 1, -1, out, &out_len, &nlp, 1) == CRNN_OK && out_len == 2 && out[0] == 1 && out[1] == 0,
          "crnn_ctc_beam_search width 2 -> [1, 0] (TensorFlow's known answer)");
    CHECK(crnn_ctc_beam_search(x, &len, 8, 1, 6, 100, 1, -1, out, &out_len, &nlp, 0) == CRNN_OK && out_len == 3 && out[0] == 0 && out[1] == 1 &&
              out[2] == 0 && out[3] == 0 && nlp > 0.f,
          "crnn_ctc_beam_search width 100 -> [0, 1, 0], zero padded");
    len = 9;
    CHECK(crnn_ctc_beam_search(x, &len, 8, 1, 6, 100, 1, -1, out, &out_len, &nlp, 1) == CRNN_INVALID_VALUE && strlen(crnn_last_error()) > 0,
          "crnn_ctc_beam_search rejects input_len > T with a status and a message");
  }

  /* host copy pool */
  {
    const size_t n = (5u << 20) + 123;
    unsigned char* a = (unsigned char*)malloc(n);
    unsigned char* b = (unsigned char*)calloc(n, 1);
    size_t i;
    for (i = 0; i < n; ++i) a[i] = (unsigned char)(i * 2654435761u >> 24);
    CHECK(a && b && crnn_host_copy(b, a, n, 4) == CRNN_OK && memcmp(a, b, n) == 0, "crnn_host_copy (4 threads)");
    CHECK(crnn_host_is_pinned(a) == 0, "crnn_host_is_pinned(malloc'd) == 0");
    free(a);
    free(b);
  }

  /* model creation: needs a CUDA device; without one it must fail with a status and a message, never fall back */
  {
    crnn_config cfg;
    crnn_model* m = NULL;
    int st;
    cfg.img_height = 32; cfg.nclasses = 64; cfg.num_hid = 512; cfg.bn_eps = 1e-3f; cfg.weight_decay = 1e-5f; cfg.compute_dtype = 1;
    st = crnn_model_create(&cfg, &m);
    if (expect_gpu) {
      CHECK(st == CRNN_OK && m != NULL && crnn_num_tensors(m) == 24 && crnn_param_count(m) == 7158592, "crnn_model_create: 24 tensors, 7 158 592 parameters");
      if (m) crnn_model_destroy(m);
    } else {
      CHECK(st != CRNN_OK && m == NULL && strlen(crnn_last_error()) > 0, "crnn_model_create without a CUDA device: status + message, no fallback");
    }
  }
  return failures ? 1 : 0;
}
