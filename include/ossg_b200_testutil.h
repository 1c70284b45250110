/*
 * include/ossg_b200_testutil.h -- C ABI of libossg_b200_testutil.so
 *
 * Measurement and test utilities that are NOT part of the reference's surface and therefore not
 * part of the product library libossg_b200.so: deterministic synthetic fields,
 * an FP64 FMA-rate probe (roofline denominator of the FP64-bound kernels), a device self-test of the
 * branch-free sqrt / division sequences, and an NVLink store-path probe.  Used by tests/, bench.py
 * and tools/ only.  Unlike the product entry points some of these allocate scratch and synchronise.
 * Same conventions otherwise: 0 on success, negative argument error, positive cudaError_t;
 * ossg_testutil_last_error() returns the thread-local message of THIS library.
 */
#ifndef OSSG_B200_TESTUTIL_H
#define OSSG_B200_TESTUTIL_H

#include "ossg_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

const char* ossg_testutil_last_error(void);

/* interior(c)[i,j,k] = hash(seed, global i, j, k) in [-1, 1); halos are filled with a NaN payload */
int ossg_synth_field(const ossg_field_desc* f, uint64_t seed, int32_t gi0, int32_t gj0, void* stream);
/* sustained FP64 FMA rate of the current device: runs `iters` dependent-chain DFMA steps per
 * thread on a full grid; returns elapsed ms via *ms and executed flops via *flops              */
int ossg_fp64_peak_probe(int iters, float* ms, double* flops, void* stream);
/* device self-test of the branch-free sqrt / division used by the tile kernel's straight-line
 * path (csrc/ossg_math.cuh sqrt_nb, div_nb) against the IEEE operators on n pseudo-random
 * operands drawn from the ranges the kernel admits; mismatches[0] = sqrt, [1] = division
 * (host pointers; the call synchronises the stream)                                            */
int ossg_selftest_sqrt_div(uint64_t n, uint64_t seed, uint64_t mismatches[2], void* stream);
/* NVLink store-path probe: kernel-issued copy of `rows` contiguous source segments of seg_bytes to
 * dst + (r / group) * group_stride + (r % group) * pitch (dst usually peer-mapped); mode 0: 8-byte
 * stores, 1: 16-byte stores, 2: shared-memory staging + one bulk store per chunk, 3: bulk load +
 * bulk store; chunk_bytes <= 16384; blocks <= 0 selects 4 blocks per SM                          */
int ossg_peer_copy_probe(void* dst, const void* src, uint64_t rows, uint32_t seg_bytes, uint64_t group,
                         uint64_t group_stride, uint64_t pitch, uint32_t chunk_bytes, int mode, int blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OSSG_B200_TESTUTIL_H */
