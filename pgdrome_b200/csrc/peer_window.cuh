// NVLink peer window shared by the sharded PCG paths (sharded.cu, pcg_persist.cu): layout of the window, the
// per-solve halo description and the system-scope release / acquire accessors.
#pragma once
#include "common.cuh"

#define PW_MAXR 16
#define PW_AR_VALS 4

struct PwLayout {
    int64_t pcap;
    __host__ __device__ size_t slot_off() const { return (size_t)pcap * 8; }
    __host__ __device__ size_t arflag_off() const { return slot_off() + 2 * PW_MAXR * PW_AR_VALS * 8; }
    __host__ __device__ size_t haloflag_off() const { return arflag_off() + 2 * PW_MAXR * 8; }
    // "LL" mailboxes of the persistent kernel's all-reduce: every double travels as two 8-byte words (32 data bits + the
    // 32-bit sequence number each), so a value is complete as soon as both words carry the expected sequence -- one
    // one-way NVLink flight, no system-scope fence and no separate flag on the critical path
    // LL halo region INSIDE the p-capacity area: doubles [llh_base, pcap) = 2 parities x llh_cap entries x 2 words; the
    // vector that lives in the window (z) keeps [0, llh_base)
    __host__ __device__ size_t llh_base() const { return ((size_t)pcap / 3) & ~(size_t)1; }
    __host__ __device__ size_t llh_cap() const { return ((size_t)pcap - llh_base()) / 4; }
    __host__ __device__ size_t ll_off() const { return haloflag_off() + PW_MAXR * 8; }
    __host__ __device__ size_t bytes() const { return ll_off() + 2 * PW_MAXR * PW_AR_VALS * 2 * 8; }
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct PwPeers {
    unsigned char* base[PW_MAXR];
};
struct PwHalo {
    int64_t seg_start[PW_MAXR + 1];  // send entries grouped by destination rank
    int64_t dst_off[PW_MAXR];        // offset (doubles) inside the destination's p where my block of ghosts starts
    int recv_from[PW_MAXR];          // 1 if I receive ghosts from that rank
};

