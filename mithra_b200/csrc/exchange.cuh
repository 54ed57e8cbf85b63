/* exchange.cuh -- z-slab exchange between the GPUs of one box (placeholder until the P2P path lands).
 *
 * Replaces the reference's blocking MPI ring (fdtd.cpp:728-799 ghost planes, fdtd.cpp:201-210 current merge,
 * solver.cpp:1552-1568 particle migration).
 */
#ifndef MITHRA_EXCHANGE_CUH_
#define MITHRA_EXCHANGE_CUH_

#include "device_types.cuh"

namespace mithra
{
  struct Exchange { bool connected; };

  static inline void exchange_init (Exchange& x) { x.connected = false; }
  static inline void exchange_destroy (Exchange&) {}
  static inline const char* exchange_error () { return "slab exchange is not built yet"; }
  static inline int exchange_export (Exchange&, const FieldDev&, double* const*, double*, float4*, void*, size_t, size_t*) { return 1; }
  static inline int exchange_connect (Exchange&, const FieldDev&, const void*, const void*) { return 1; }
  static inline int exchange_potentials (Exchange&, const FieldDev&, double*, cudaStream_t) { return 1; }
  static inline int exchange_eb (Exchange&, const FieldDev&, float4*, cudaStream_t) { return 1; }
  static inline int exchange_current (Exchange&, const FieldDev&, double*, Box*, cudaStream_t) { return 1; }
}

#endif
