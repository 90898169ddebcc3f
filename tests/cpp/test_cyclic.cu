// Host-only check of the tile-cyclic distribution helpers of csrc/pso.cuh (sequential-exact ParticleSwarm over several
// ranks): the ranks' index sets partition 0..N-1, local order is global order, the three index maps agree with a brute-force
// walk.  Compiled with nvcc because the header is CUDA; nothing here touches a device.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../eqsolver_b200/csrc/pso.cuh"

using namespace eqs;

static int fail(const char *what, long long N, int world, int rank, long long i)
{
    std::printf("FAIL %s: N=%lld world=%d rank=%d at %lld\n", what, N, world, rank, i);
    return 1;
}

int main()
{
    const long long sizes[] = {0, 1, 2, 255, 256, 257, 511, 512, 513, 1000, 2003, 4096, 70001, 262144 + 77};
    long long checked = 0;
    for (long long N : sizes) {
        for (int world = 1; world <= 9; ++world) {
            std::vector<int> seen((size_t)N, 0);
            long long total = 0;
            for (int rank = 0; rank < world; ++rank) {
                long long nloc = 0, per = 0;
                pso_cyclic_shard(N, world, rank, &nloc, &per);
                PsoDev P{};
                P.cyclic = 1;
                P.world = world;
                P.rank = rank;
                P.nloc = nloc;
                P.ntotal = N;
                const long long ntiles = (N + PSO_GROUP - 1) / PSO_GROUP;
                if (per * world < ntiles || per < 1) return fail("tiles_per_rank", N, world, rank, per);
                long long prev = -1;
                for (long long li = 0; li < nloc; ++li) {
                    const long long g = pso_gidx(P, li);
                    if (g < 0 || g >= N) return fail("gidx range", N, world, rank, li);
                    if (g <= prev) return fail("local order is not global order", N, world, rank, li);
                    if ((g / PSO_GROUP) % world != rank) return fail("owner", N, world, rank, li);
                    if (pso_lidx(P, g) != li) return fail("lidx(gidx)", N, world, rank, li);
                    if (pso_lcount(P, g) != li) return fail("lcount(gidx)", N, world, rank, li);         // owned particles below g
                    if (pso_lcount(P, g + 1) != li + 1) return fail("lcount(gidx + 1)", N, world, rank, li);
                    if ((g / PSO_GROUP) / world >= per) return fail("local tile beyond tiles_per_rank", N, world, rank, li);
                    seen[(size_t)g]++;
                    prev = g;
                    ++checked;
                }
                // lcount at arbitrary window bounds: brute force over a few cuts
                for (long long cut : {0LL, 1LL, 100LL, 256LL, 300LL, N / 3, N / 2 + 7, N - 1, N, N + 5}) {
                    if (cut < 0) continue;
                    long long want = 0;
                    for (long long li = 0; li < nloc; ++li) want += pso_gidx(P, li) < cut;
                    if (pso_lcount(P, cut) != want) return fail("lcount(cut)", N, world, rank, cut);
                }
                total += nloc;
            }
            if (total != N) return fail("sizes do not add up", N, world, -1, total);
            for (long long g = 0; g < N; ++g)
                if (seen[(size_t)g] != 1) return fail("not a partition", N, world, -1, g);
        }
    }
    // the contiguous distribution through the same helpers
    PsoDev P{};
    P.first = 1000;
    P.nloc = 50;
    if (pso_gidx(P, 7) != 1007 || pso_lidx(P, 1007) != 7 || pso_lcount(P, 990) != 0 || pso_lcount(P, 1010) != 10 ||
        pso_lcount(P, 5000) != 50)
        return fail("contiguous", 0, 1, 0, 0);
    std::printf("cyclic distribution ok (%lld index checks)\n", checked);
    return 0;
}
