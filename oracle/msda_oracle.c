/*
 * TEST INFRASTRUCTURE ONLY.  CPU oracle for the multi-scale deformable attention hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library. The product (uninext_b200/) never links or calls it; the product fails loudly without its CUDA
 * library instead of falling back to this code.
 *
 * What it restates: ms_deformable_im2col_gpu_kernel / ms_deformable_col2im_gpu_kernel_* and their bilinear
 * helpers in /root/reference/projects/UNINEXT/uninext/models/deformable_detr/ops/src/cuda/
 * ms_deform_im2col_cuda.cuh (line citations inside msda_oracle_impl.h). The reference has no CPU kernel of its
 * own (ops/src/cpu/ms_deform_attn_cpu.cpp:17-41 only raises), so this file is a "port", not the reference.
 *
 * Parity pin: tests/test_oracle_golden.py checks these functions against the .npz files under tests/golden, which were
 * produced in the build container by importing the reference's own ms_deform_attn_core_pytorch
 * (ops/functions/ms_deform_attn_func.py:43-63) and differentiating it with torch.autograd in fp64
 * (generator: tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -fopenmp -fPIC -shared -o oracle/libmsda_oracle.so oracle/msda_oracle.c -lm   (oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define REAL double
#define FN(name) name##_f64
#include "msda_oracle_impl.h"
#undef REAL
#undef FN

#define REAL float
#define FN(name) name##_f32
#include "msda_oracle_impl.h"
#undef REAL
#undef FN

int msda_oracle_abi_version(void) { return 1; }
