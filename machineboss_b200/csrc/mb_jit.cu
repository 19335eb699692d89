// mb_jit.cu -- placeholder until the NVRTC-specialised strip engine lands.
#include "mb_internal.h"
namespace mb {
bool jit_supported (const mb_machine*, std::string* why) { if (why) *why = "jit engine not built"; return false; }
int jit_prepare (mb_machine*) { set_error ("jit engine not built"); return 1; }
void jit_destroy (mb_machine*) { }
int jit_update_weights (mb_machine*) { return 0; }
int jit_forward (mb_machine*, mb_batch*, double*, bool) { set_error ("jit engine not built"); return 1; }
int jit_viterbi (mb_machine*, mb_batch*, double*, int64_t*) { set_error ("jit engine not built"); return 1; }
int jit_counts (mb_machine*, mb_batch*, double*, double*) { set_error ("jit engine not built"); return 1; }
}
