// fc_async.cuh — named-barrier helper of the fused axis kernel (a barrier over the compute warps only).
// tests/cpu_emul supplies a host stand-in with the same semantics. The tensor-core GEMM keeps its own mbarrier /
// bulk-copy / tcgen05 wrappers in fc_tc.cuh (they have no host emulation).
#pragma once
#include "fc_kernels.cuh"

#ifdef FC_CPU_EMUL
void fc_emul_named_barrier(int id, int count);  // tests/cpu_emul/cuda_shim.cpp
FC_DEV void fc_named_bar_sync(int id, int threads) { fc_emul_named_barrier(id, threads); }
#else
FC_DEV void fc_named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
#endif
