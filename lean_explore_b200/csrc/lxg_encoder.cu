// Sentence-encoder entry points (include/lxg.h).  Placeholder until the BERT kernels land.
#include <string>

#include "../../include/lxg.h"
#include "common.h"

extern "C" {
int lxg_encoder_create(lxg_encoder** out, const lxg_bert_weights*) {
  if (out) *out = nullptr;
  return lxg::set_error(LXG_EUNSUPPORTED, "encoder kernels not built yet");
}
int lxg_encoder_destroy(lxg_encoder*) { return LXG_OK; }
int lxg_encode(lxg_encoder*, const int32_t*, const int32_t*, int32_t, int32_t, int, float*, void*) {
  return lxg::set_error(LXG_EUNSUPPORTED, "encoder kernels not built yet");
}
}
