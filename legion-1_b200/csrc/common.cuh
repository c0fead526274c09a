// common.cuh -- shared device helpers for the sm_100a Legion hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/legion_b200.h"

// Debug build (`make -C legion-1_b200/csrc debug` -> _build/liblegion_b200_dbg.so, -DLGN_DEBUG): a device-side assert
// in front of every indexed write of the sampling kernels.  compute-sanitizer is closed on the measurement pool;
// the parity suite run against this build (LGN_LIBRARY=<path> pytest -m gpu) is the out-of-bounds check.
#ifdef LGN_DEBUG
#include <assert.h>
#define LGN_ASSERT(c) assert(c)
#else
#define LGN_ASSERT(c) ((void)0)
#endif

namespace lgn {

// ---- dedup value encoding ------------------------------------------------
// One 32-bit value per node of the batch replaces the reference's dedup bitmap + position_map
// (Kernels.cu:88-92, 412-438):
//     [31] 0   [30:25] generation   [24] CAND   [23:0] payload
// payload without CAND : final local index of the node
// payload with CAND    : smallest slot of the current hop that sampled the node (a candidate)
// EMPTY                : generation 63, never written by a batch
// Every value a batch writes carries the batch's generation, and generations DEcrease from batch to batch
// (62, 61, .., 0, then the table is reset): a red.min therefore overwrites whatever an older batch left
// behind, and nothing has to be released at the end of a batch (the reference memsets N/8 bytes per batch and
// clears its position map entry by entry, Kernels.cu:750-756).
constexpr int32_t EMPTY = 0x7fffffff;
constexpr int GEN_SHIFT = 25;
constexpr int N_GEN = 63;
constexpr int32_t CAND = 1 << 24;
constexpr int32_t PAYLOAD_MASK = (1 << GEN_SHIFT) - 1;

// ---- per-batch device state (written by kernels, never read by the host on
// the hot path) ----------------------------------------------------------
struct HopState {
    int32_t n_items;     // frontier size F of this hop (nc[2])
    int32_t item_base;   // offset of the frontier inside agg_src_ids (ec[2]); hop 0: seeds
    int32_t node_base;   // write cursor into sampled_ids (nc[0])
    int32_t edge_base;   // write cursor into the agg arrays (ec[0])
};

struct BatchState {
    HopState hop[LGN_MAX_HOPS + 1];
    uint32_t step;          // philox counter word 3 (batch id inside the epoch, + the mode's offset)
    int32_t status;         // sticky: 0 or LGN_E_CAPACITY
    int32_t max_ids;        // max unique ids over presampled batches (GPUCache.cu:294-296)
    uint32_t epoch;         // philox counter word 1
    int32_t gen_base;       // generation of this batch << GEN_SHIFT (dedup values, see above)
    uint32_t pad1;
    int32_t dbg_max_slots;  // bounds the debug build asserts against (slots of the widest hop, id / edge capacity)
    int32_t dbg_capacity;
    unsigned long long tier_rows[4];   // local, peer, host rows gathered
    unsigned long long tot_items;      // frontier items expanded since the last reset (every hop)
    unsigned long long tot_edges;      // edges sampled since the last reset
};

// ---- cache-hinted 128-bit / 32-bit accesses (G13/G14 of the Blackwell guide) ----
__device__ __forceinline__ uint4 ld_nc_v4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ld_nc_u32(const void* p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_cs_v4(void* p, uint4 v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_cs_u32(void* p, uint32_t v)
{
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- L2 residency control ---------------------------------------
// The dedup map is touched at random ~5 times per sampled edge, the feature rows exactly once:
// slot_map accesses carry an evict_last policy so the map stays in the 126 MB L2 while the
// gathers stream through it with evict_first.
__device__ __forceinline__ unsigned long long policy_evict_last()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long policy_evict_first()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void red_min_keep(int32_t* p, int32_t v, unsigned long long pol)
{
    asm volatile("red.global.min.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ int32_t ld_keep(const int32_t* p, unsigned long long pol)
{
    int32_t r;
    asm volatile("ld.global.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ void st_keep(int32_t* p, int32_t v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.s32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint4 ld_stream_v4(const void* p, unsigned long long pol)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_stream_v4(void* p, uint4 v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

// ---- per-batch dedup structure ------------------------------------------------------------
// Two interchangeable layouts behind one interface:
//   direct  int32 map[N]: one probe per access, but 4N bytes per lane: L2-resident only for small graphs.
//   hash    open-addressing table of (key << 32 | value) words sized for the BATCH (2^bits entries, a few MB),
//           so it stays in L2 whatever N is.  A 64-bit red.min keeps the smallest value of a key because the key
//           occupies the high word.  A claim returns the entry's index ("handle"); later passes address the
//           entry directly, without probing.  Entries of older generations count as free.
constexpr unsigned long long EMPTY64 = ~0ull;

struct Dedup {
    int32_t* map;
    unsigned long long* tab;
    uint32_t bits;            // 0 = direct map
};

__device__ __forceinline__ unsigned long long ld_keep_u64(const unsigned long long* p, unsigned long long pol)
{
    unsigned long long r;
    asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ void st_keep_u64(unsigned long long* p, unsigned long long v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_min_keep_u64(unsigned long long* p, unsigned long long v, unsigned long long pol)
{
    asm volatile("red.global.min.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}

// lower the value stored for `key` to `val` (generation | payload); returns the entry's handle, or -1 if the hash
// table is full (reported through BatchState::status)
__device__ __forceinline__ int32_t dedup_claim(const Dedup& dd, int32_t key, int32_t val, unsigned long long keep)
{
    if (dd.bits == 0) { red_min_keep(&dd.map[key], val, keep); return key; }
    const uint32_t mask = (1u << dd.bits) - 1u;
    uint32_t h = (uint32_t)key;          // murmur3 finaliser: node ids are structured (hub scatter, strided seed lists),
    h ^= h >> 16; h *= 0x85ebca6bu;      // a single multiplicative hash clusters badly on them
    h ^= h >> 13; h *= 0xc2b2ae35u;
    h ^= h >> 16;
    h &= mask;
    const uint32_t gen = (uint32_t)val >> GEN_SHIFT;
    const unsigned long long mine = ((unsigned long long)(uint32_t)key << 32) | (uint32_t)val;
    for (int probe = 0; probe < 1024; probe++) {
        unsigned long long cur = ld_keep_u64(dd.tab + h, keep);
        if (((uint32_t)cur >> GEN_SHIFT) != gen) {                 // free: never used (EMPTY64 is generation 63) or left by an older batch
            const unsigned long long seen = atomicCAS(dd.tab + h, cur, mine);   // ptxas rejects .L2::cache_hint on atom.cas
            if (seen == cur) return (int32_t)h;
            cur = seen;                                              // somebody of this batch took it: fall through and look at it
            if (((uint32_t)cur >> GEN_SHIFT) != gen) { probe--; continue; }
        }
        if ((uint32_t)(cur >> 32) == (uint32_t)key) {
            if ((uint32_t)cur > (uint32_t)val) red_min_keep_u64(dd.tab + h, mine, keep);   // values only ever decrease
            return (int32_t)h;
        }
        h = (h + 1u) & mask;
    }
    return -1;
}
// payload (CAND | slot, or final index) of an entry claimed in this batch
__device__ __forceinline__ int32_t dedup_payload(const Dedup& dd, int32_t handle, unsigned long long keep)
{
    const int32_t v = dd.bits == 0 ? ld_keep(&dd.map[handle], keep) : (int32_t)(uint32_t)ld_keep_u64(dd.tab + handle, keep);
    return v & PAYLOAD_MASK;
}
// payload and key in one access (the hash entry holds both; a direct-map handle IS the key)
__device__ __forceinline__ int32_t dedup_payload_key(const Dedup& dd, int32_t handle, int32_t& key, unsigned long long keep)
{
    if (dd.bits == 0) { key = handle; return ld_keep(&dd.map[handle], keep) & PAYLOAD_MASK; }
    const unsigned long long e = ld_keep_u64(dd.tab + handle, keep);
    key = (int32_t)(uint32_t)(e >> 32);
    return (int32_t)(uint32_t)e & PAYLOAD_MASK;
}
__device__ __forceinline__ void dedup_publish(const Dedup& dd, int32_t handle, int32_t key, int32_t val, unsigned long long keep)
{
    if (dd.bits == 0) st_keep(&dd.map[handle], val, keep);
    else st_keep_u64(dd.tab + handle, ((unsigned long long)(uint32_t)key << 32) | (uint32_t)val, keep);
}

// ---- thrust::minstd_rand compatibility (Kernels.cu:402-405) ----------------
// state after discard(z) from seed 1 is 48271^z mod (2^31-1); the next draw is
// 48271^(z+1).  Mersenne modulus => fold instead of divide.
__device__ __forceinline__ uint32_t mulmod_m31(uint32_t a, uint32_t b)
{
    unsigned long long p = (unsigned long long)a * b;          // < 2^62
    unsigned long long f = (p & 0x7fffffffull) + (p >> 31);     // < 2^32
    uint32_t r = (uint32_t)(f & 0x7fffffffull) + (uint32_t)(f >> 31);
    return r >= 0x7fffffffu ? r - 0x7fffffffu : r;
}
__device__ __forceinline__ uint32_t minstd_pow(unsigned long long e)
{
    uint32_t base = 48271u, acc = 1u;
    while (e) {
        if (e & 1ull) acc = mulmod_m31(acc, base);
        e >>= 1;
        base = mulmod_m31(base, base);
    }
    return acc;
}
// uniform_int_distribution<int>(0, deg-1) on top of the raw draw x (thrust
// uniform_real_distribution.inl: double(x-min)/(1+double(max-min)) * deg).
__device__ __forceinline__ int32_t minstd_to_pick(uint32_t x, int32_t deg)
{
    double u = (double)(x - 1u);
    u = __ddiv_rn(u, 2147483646.0);
    return (int32_t)__dmul_rn(u, (double)deg);
}

// ---- Philox4x32-10, counter = (slot, epoch, hop, step), key = seed ----
// A hop's slot index is < 2^30 (lgn_create bounds the capacity by CAND), so counter word 1 carries the epoch:
// every (seed, epoch, step, hop, slot) draws from its own block; epoch 0 is the stream of round 1.
__device__ __forceinline__ uint32_t philox_first_word(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                      uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}
__device__ __forceinline__ int32_t philox_pick(unsigned long long idx, uint32_t epoch, uint32_t hop, uint32_t step,
                                               unsigned long long seed, int32_t deg)
{
    uint32_t r = philox_first_word((uint32_t)idx, epoch + (uint32_t)(idx >> 32), hop, step, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (int32_t)__umulhi(r, (uint32_t)deg);
}

}  // namespace lgn
