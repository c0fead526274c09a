// Pin for the oracle's minstd closed form: runs the CUDA toolkit's own thrust
// headers on the HOST (THRUST_DEVICE_SYSTEM_CPP) through the exact call sequence
// of the reference (Kernels.cu:402-405) and prints a JSON fixture.
// Build/run: see oracle/Makefile target `pins` (writes tests/golden/minstd_pin.json).
#include <thrust/random/linear_congruential_engine.h>
#include <thrust/random/uniform_int_distribution.h>
#include <cstdio>
#include <cstdint>
#include <vector>

static int pick(unsigned long long idx, int deg)
{
    thrust::minstd_rand engine;
    engine.discard(idx);
    thrust::uniform_int_distribution<> dist(0, deg - 1);
    return dist(engine);
}

int main()
{
    const int degs[] = {1, 2, 3, 7, 15, 25, 26, 100, 1000, 100000, 2000000000};
    std::vector<unsigned long long> idxs;
    for (unsigned long long i = 0; i < 64; i++) idxs.push_back(i);
    unsigned long long x = 12345;
    for (int i = 0; i < 192; i++) { x = x * 6364136223846793005ull + 1442695040888963407ull; idxs.push_back((x >> 33) % 20000000ull); }
    idxs.push_back(2147483645ull); idxs.push_back(2147483646ull); idxs.push_back(2147483647ull); idxs.push_back(4000000000ull);
    printf("{\n \"source\": \"thrust (CUDA toolkit) minstd_rand().discard(idx); uniform_int_distribution<>(0,deg-1)\",\n");
    thrust::minstd_rand e0; printf(" \"first_raw\": %u,\n", (unsigned)e0());
    printf(" \"degs\": [");
    for (size_t d = 0; d < sizeof(degs)/sizeof(int); d++) printf("%s%d", d ? ", " : "", degs[d]);
    printf("],\n \"idx\": [");
    for (size_t i = 0; i < idxs.size(); i++) printf("%s%llu", i ? ", " : "", idxs[i]);
    printf("],\n \"picks\": [\n");
    for (size_t d = 0; d < sizeof(degs)/sizeof(int); d++) {
        printf("  [");
        for (size_t i = 0; i < idxs.size(); i++) printf("%s%d", i ? ", " : "", pick(idxs[i], degs[d]));
        printf("]%s\n", d + 1 < sizeof(degs)/sizeof(int) ? "," : "");
    }
    printf(" ]\n}\n");
    return 0;
}
