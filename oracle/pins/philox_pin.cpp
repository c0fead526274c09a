// Pin for the oracle's Philox4x32-10: libcu++'s cuda::std::philox4x32 (C++26
// std::philox4x32: n=4, w=32, r=10, multipliers 0xD2511F53/0xCD9E8D57, round
// constants 0x9E3779B9/0xBB67AE85) run on the HOST, printed as a JSON fixture.
// Engine output order is X[0..3] of consecutive counters starting at counter+1?
// -- no: the engine emits the block of the CURRENT counter then increments, so
// with set_counter({c3,c2,c1,c0}) (most-significant word first) the first four
// outputs are the block of (c0,c1,c2,c3).
#include <cuda/std/__random/philox_engine.h>
#include <cstdio>
#include <cstdint>
struct two_word_seq {
    uint32_t k0, k1;
    using result_type = uint32_t;
    template <class It> void generate(It b, It e) { uint32_t v[2] = {k0, k1}; int i = 0; for (; b != e; ++b, ++i) *b = v[i % 2]; }
};
int main()
{
    struct { uint32_t c[4]; uint32_t k[2]; } cases[] = {
        {{0, 0, 0, 0}, {0, 0}},
        {{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu}},
        {{0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u}, {0xa4093822u, 0x299f31d0u}},
        {{7, 0, 1, 42}, {0x1E6104u, 0}},
        {{123456789u, 1, 2, 1000}, {0xdeadbeefu, 0xcafef00du}},
    };
    printf("{\n \"source\": \"libcu++ cuda::std::philox4x32 on host\",\n \"cases\": [\n");
    for (size_t i = 0; i < sizeof(cases)/sizeof(cases[0]); i++) {
        cuda::std::philox4x32 e;
        // seed() sets key word 0 only; set full key through seed sequence is awkward -> use
        // the engine's public API: seed(value) => K[0]=value.  For two-word keys fall back
        // to constructing with a seed_seq-like object below.
        two_word_seq seq{cases[i].k[0], cases[i].k[1]};
        e.seed(seq);
        e.set_counter({cases[i].c[3], cases[i].c[2], cases[i].c[1], cases[i].c[0]});
        uint32_t o[4]; for (int j = 0; j < 4; j++) o[j] = (uint32_t)e();
        printf("  {\"ctr\": [%u, %u, %u, %u], \"key\": [%u, %u], \"out\": [%u, %u, %u, %u]}%s\n",
               cases[i].c[0], cases[i].c[1], cases[i].c[2], cases[i].c[3], cases[i].k[0], cases[i].k[1],
               o[0], o[1], o[2], o[3], i + 1 < sizeof(cases)/sizeof(cases[0]) ? "," : "");
    }
    cuda::std::philox4x32 d; uint32_t v = 0; for (int i = 0; i < 10000; i++) v = (uint32_t)d();
    printf(" ],\n \"default_10000th\": %u\n}\n", v);
    return 0;
}
