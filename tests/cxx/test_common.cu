// Host-side run of the __host__ __device__ helpers of hysortk_b200/csrc/common.cuh (the same source the kernels
// compile): for every "<k> <w0> <w1> <w2>" line on stdin prints the reverse complement and the canonical form, and for
// every "slot <sw> <len> <hex words...>" nothing else yet.  The Python test compares with the oracle's restatement of the
// reference (include/kmer.hpp:265-303).
#include <cstdio>
#include <cstdlib>
#include "../../hysortk_b200/csrc/common.cuh"

using namespace hsk;

template <int NW>
static void one(int k, const unsigned long long *in)
{
    u64 w[NW], t[NW];
    for (int l = 0; l < NW; ++l) w[l] = in[l];
    kmer_twin<NW>(w, k, t);
    kmer_canonical<NW>(w, k);
    for (int l = 0; l < 3; ++l) std::printf("%llx ", l < NW ? t[l] : 0ull);
    for (int l = 0; l < 3; ++l) std::printf("%llx ", l < NW ? w[l] : 0ull);
    std::printf("\n");
}

int main()
{
    int k;
    unsigned long long in[3];
    while (std::scanf("%d %llx %llx %llx", &k, &in[0], &in[1], &in[2]) == 4) {
        const int nw = nwords_for_k(k);
        if (nw == 1) one<1>(k, in); else if (nw == 2) one<2>(k, in); else one<3>(k, in);
    }
    return 0;
}
