// ctx_internal.h -- what comm.cu needs from a smash_ctx (the struct itself stays private to api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/smash_b200.h"

namespace smash { struct TailState; }

int ctx_fail(int code, const char *fmt, ...);                  // sets smash_last_error() of the calling thread, returns code
void **ctx_comm_slot(smash_ctx *c);                            // the context's communicator state (owned by comm.cu)
int ctx_device(const smash_ctx *c);
smash::TailState *ctx_tail(smash_ctx *c);
cudaStream_t ctx_stream(smash_ctx *c);
uint64_t *ctx_launches(smash_ctx *c);
uint64_t ctx_n_bins(smash_ctx *c);
int ctx_require_idle(smash_ctx *c);                            // SMASH_ERR_STATE while a slot has a batch in flight
