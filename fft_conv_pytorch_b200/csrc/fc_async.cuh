// fc_async.cuh — mbarrier, bulk async copy (cp.async.bulk, the TMA engine without a tensor map) and named-barrier
// helpers shared by the fused axis kernel and the tensor-core GEMM. tests/cpu_emul supplies host stand-ins with the
// same semantics (a phase completes when all expected arrivals and all expected bytes have been counted).
#pragma once
#include "fc_kernels.cuh"

#ifdef FC_CPU_EMUL
// provided by tests/cpu_emul/cuda_shim.*
struct fc_mbar {
  void* impl;
};
void fc_emul_mbar_init(fc_mbar* b, unsigned count);
void fc_emul_mbar_arrive(fc_mbar* b);
void fc_emul_mbar_expect_tx(fc_mbar* b, unsigned bytes);
void fc_emul_mbar_complete_tx(fc_mbar* b, unsigned bytes);
void fc_emul_mbar_wait(fc_mbar* b, unsigned parity);
void fc_emul_named_barrier(int id, int count);
FC_DEV void fc_mbar_init(fc_mbar* b, unsigned count) { fc_emul_mbar_init(b, count); }
FC_DEV void fc_mbar_init_fence() {}
FC_DEV void fc_mbar_arrive(fc_mbar* b) { fc_emul_mbar_arrive(b); }
FC_DEV void fc_mbar_expect_tx(fc_mbar* b, unsigned bytes) { fc_emul_mbar_expect_tx(b, bytes); }
FC_DEV void fc_mbar_wait(fc_mbar* b, unsigned parity) { fc_emul_mbar_wait(b, parity); }
FC_DEV void fc_bulk_g2s(void* dst, const void* src, unsigned bytes, fc_mbar* b) {
  std::memcpy(dst, src, bytes);
  fc_emul_mbar_complete_tx(b, bytes);
}
FC_DEV void fc_named_bar_sync(int id, int threads) { fc_emul_named_barrier(id, threads); }
#else
struct fc_mbar {
  unsigned long long v;
};
FC_DEV unsigned fc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
FC_DEV void fc_mbar_init(fc_mbar* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fc_smem_u32(b)), "r"(count));
}
FC_DEV void fc_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
FC_DEV void fc_mbar_arrive(fc_mbar* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fc_smem_u32(b)) : "memory"); }
FC_DEV void fc_mbar_expect_tx(fc_mbar* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fc_smem_u32(b)), "r"(bytes) : "memory");
}
FC_DEV void fc_mbar_wait(fc_mbar* b, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(fc_smem_u32(b)),
      "r"(parity)
      : "memory");
}
// 1-d bulk async copy global -> shared; completion is counted in bytes on the mbarrier. 16-byte aligned, size % 16 == 0.
FC_DEV void fc_bulk_g2s(void* dst, const void* src, unsigned bytes, fc_mbar* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(fc_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(fc_smem_u32(b))
               : "memory");
}
FC_DEV void fc_named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
#endif
