// smz_net_bf16.cu — placeholder until the tcgen05 path lands: creating a BF16 engine fails loudly.
#include <stdio.h>

#include "../../include/smz.h"
#include "smz_net_bf16.h"

struct SmzBf16Image { int unused; };

int smz_bf16_create(const SmzNetShape&, const SmzArena&, SmzBf16Image**, char* err, size_t err_len) {
  snprintf(err, err_len, "SMZ_NET_BF16 is not available in this build");
  return SMZ_E_STATE;
}
void smz_bf16_destroy(SmzBf16Image*) {}
int smz_bf16_pack(SmzBf16Image*, const SmzNetShape&, const float*, cudaStream_t, char*, size_t) { return SMZ_E_STATE; }
void smz_bf16_root(SmzBf16Image*, const SmzArena&, const SmzNetShape&, int, const float*, cudaStream_t) {}
void smz_bf16_sim(SmzBf16Image*, const SmzArena&, const SmzNetShape&, int, int, cudaStream_t) {}
void smz_bf16_eval(SmzBf16Image*, const SmzNetShape&, int, int, const float*, const int*, float*, float*, float*,
                   float*, int*, int, cudaStream_t) {}
