// Test infrastructure: the whole image front-end (leopard_b200/csrc/lp_frontend.cu — kernels AND the host function that sequences
// them, with the vertical-first rule, the pad geometry and the workspace carving) run on the CPU.  `frontend_src.inc` is the
// file itself with its three #include lines removed, the dynamic shared-memory declaration dropped (the buffer below) and its
// launches rewritten (tests/emu/launch_rewrite.py).  Exported: lp_image_frontend / lp_image_frontend_workspace_bytes, unchanged.
#include "host_emu.h"

namespace lp {
static uint4 fe_smem[16384];   // 256 KB: the kernels' dynamic shared memory (blocks run one after another)
}
#include "elem_helpers.inc"
#include "frontend_src.inc"
