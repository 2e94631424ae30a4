// Shared host/device helpers for libcrnnctc.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/crnn_ctc.h"
#include "ptx.cuh"

int crnn_fail(int status, const char* fmt, ...);   // records crnn_last_error(), returns status

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return crnn_fail(CRNN_CUDA_ERROR, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define CRNN_TRY(expr)            \
  do {                            \
    int _s = (expr);              \
    if (_s != CRNN_OK) return _s; \
  } while (0)
