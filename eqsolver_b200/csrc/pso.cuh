// pso.cuh -- device-side state of one ParticleSwarm shard and the host driver class.
//
// Data layout in HBM: array of structures of arrays, 4 dimensions of one particle per 32-byte sector
//   x, v, pbest : [ceil(n/4)][ld][4] f64   (ld = local particle count rounded up to 32)
//   pbest_cost  : [ld] f64
// replacing the reference's array of structs `Vec<Particle>` (particle_swarm.rs:14-22, 361-389): lane i of a warp
// moves group q of particle i with one 256-bit access, a warp touches 1 KB of consecutive bytes.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace eqs {

// device-resident control block: everything the outer loop (particle_swarm.rs:394-428) decides on
struct PsoCtrl {
    double Gc;                      // global_best_cost
    double ring[EQS_STALL_WINDOW];  // CircularArray<T,50> previous_best (:24-46,359)
    long long ring_i;
    long long iter;                 // outer-loop passes that ran
    long long improvements;         // gbest improvements so far
    long long spec_first;           // sequential mode: first improving global index of the last round, -1 = none
    int stalled;                    // result of stalled_too_long (:433-441) for the NEXT pass
    int p2p_timeout;                // a peer never showed up in the fused exchange
    int pass_done;                  // sequential mode: the speculate-and-repair pass is complete
    unsigned int block_counter;     // last-block-done ticket
    int nan_seen;
    // sequential mode, windowed device-driven rounds (pso_seqw_kernel)
    long long seq_start;            // first particle of the current pass that is not final yet
    long long seq_commit_lo, seq_commit_hi; // particles whose pbest update (:417-419) the next launch performs
    long long seq_window;           // particles speculated per round, re-chosen after every pass (pso_seqw_kernel)
    long long seq_wmin, seq_wmax;
    long long seq_imp0;             // `improvements` when the current pass began
    double seq_kest;                // running estimate of gbest improvements per pass
    double seq_wscale;              // 2 * (fixed cost of a round) / (cost of one particle), in particles
    long long seq_target;           // launches are no-ops once iter reaches it
    int seq_parity;                 // 0: the committed (x, v) live in x / v, 1: in x2 / v2
    int seq_commit_buf;             // which of x (0) / x2 (1) holds the positions of the pending commit
    unsigned long long seq_round;   // rounds so far (picks the slot below)
    unsigned long long seq_slot[3]; // first improving particle of round r in slot r % 3 (atomicMin; ~0 = none)
    unsigned int seq_barrier;       // arrivals at the grid barrier of the current launch of pso_seqp_kernel
    unsigned long long seq_xdone;   // several ranks: rounds whose cross-rank result CTA 0 has published in recv (monotonic)
};

struct PsoDev {
    double *x, *v, *pb, *pbc;       // population, [ceil(n/4)][ld][4]; synchronous mode: x = buffer A, pb = buffer B (pso.cu)
    double *cost;                   // [ld] costs of the last evaluation (init ring scan, sequential mode)
    unsigned char *pbflag;          // [ld] synchronous mode: bit 0 = stale (pbest == position), bit 1 = x lives in B
    double *x2, *v2;                // sequential mode: second (x, v) pair -- speculated state / role-swapping partner
    double *G;                      // [n] global_best_position
    PsoCtrl *ctrl;
    double *part_cost;              // [grid] per-block partial min-loc
    long long *part_idx;
    double *group_min;              // [tiles per rank] per-tile minimum of the init costs
    double *gm_all;                 // cyclic distribution: every rank's group_min, [world][tiles_per_rank]
    long long tiles_per_rank;       // cyclic distribution: ceil(ceil(ntotal / GROUP) / world)
    double *send, *recv;            // exchange records: [R], [world][R]
    char *mailbox;                  // peer-memory exchange (common.cuh): this rank's mailbox, or nullptr
    char *const *peers;             // device table of every rank's mailbox as mapped on this GPU
    long long p2p_timeout_ns;       // how long the fused exchange waits for a peer (comm_p2p_timeout_ns)
    const double *lower, *upper;    // [n]
    const double *replay;           // device copy of the caller's draw stream, or nullptr
    long long ld, nloc, first, ntotal;
    long long replay_t0;            // iteration index of replay block 0
    int n, world, rank;
    int sync;                       // synchronous mode: role-swapping buffers, flags in pbflag
    // Which particles a rank holds.  0: the contiguous block [first, first + nloc) (synchronous mode, single rank).
    // 1: TILE-CYCLIC -- global tile T (PSO_GROUP consecutive particles) lives on rank T % world as local tile T / world.
    // Sequential-exact mode over several ranks needs it: the reference's order is the global index order, a window of
    // that order must keep EVERY rank busy, and with contiguous blocks it would lie inside one rank's block.
    int cyclic;
    double w, c1, c2, tol;
    PhiloxKey key;
};

constexpr int PSO_BLOCK = 256;
constexpr int PSO_GROUP = 256; // particles per init tile = granularity of the ring scan and of the cyclic distribution
constexpr int PSO_GROUP_LOG2 = 8;
static_assert((1 << PSO_GROUP_LOG2) == PSO_GROUP, "tile size");

// global index of local particle li / local index of owned global particle g / owned particles with global index < g
__host__ __device__ inline long long pso_gidx(const PsoDev &P, long long li)
{
    return P.cyclic ? ((((li >> PSO_GROUP_LOG2) * P.world + P.rank) << PSO_GROUP_LOG2) | (li & (PSO_GROUP - 1))) : P.first + li;
}
__host__ __device__ inline long long pso_lidx(const PsoDev &P, long long g)
{
    return P.cyclic ? ((((g >> PSO_GROUP_LOG2) / P.world) << PSO_GROUP_LOG2) | (g & (PSO_GROUP - 1))) : g - P.first;
}
__host__ __device__ inline long long pso_lcount(const PsoDev &P, long long g)
{
    long long c;
    if (P.cyclic) {
        const long long row = (long long)PSO_GROUP * P.world, r = g / row, rem = g % row, t = rem >> PSO_GROUP_LOG2;
        c = r * PSO_GROUP + (P.rank < t ? PSO_GROUP : (P.rank == t ? (rem & (PSO_GROUP - 1)) : 0));
    } else {
        c = g - P.first;
    }
    return c < 0 ? 0 : (c > P.nloc ? P.nloc : c);
}
// tile-cyclic distribution, host side: how many of `count` particles rank `rank` of `world` holds, and how many tiles a
// rank holds at most (the stride of the gathered tile minima)
inline void pso_cyclic_shard(long long count, int world, int rank, long long *nloc, long long *tiles_per_rank)
{
    const long long ntiles = (count + PSO_GROUP - 1) / PSO_GROUP;
    const long long owned = ntiles > rank ? (ntiles - rank + world - 1) / world : 0;
    long long n = owned * PSO_GROUP;
    if (owned > 0 && (ntiles - 1) % world == rank) n -= ntiles * PSO_GROUP - count; // the ragged last tile
    *nloc = n;
    const long long per = (ntiles + world - 1) / world;
    *tiles_per_rank = per > 1 ? per : 1;
}
constexpr int PSO_SMALL_BLOCK = 512; // K8: threads of the single-CTA whole-solve kernel
constexpr int PSO_SMALL_MAX = 2048;  // K8: largest population it takes (4 particles per thread)

int validate_pso_params(const eqs_pso_params *p, std::string *why);

class PsoSolver {
public:
    PsoSolver(Ctx *ctx) : ctx_(ctx) {}
    ~PsoSolver();
    int create(const eqs_pso_params *p);
    int init(const double *x0);
    int init_local(const double *x0, int phase);
    int init_apply();
    int step(int64_t iters, const double *replay_step);
    int step_local(const double *replay_step);
    int step_apply();
    int sync();
    int read_gbest(double *g, double *cost, int64_t *iters, int32_t *stalled, int64_t *improvements);
    int read_state(double *x, double *v, double *pb, double *pbc);
    int read_ring(double *ring, int64_t *ring_i);
    int local_indices(int64_t *gidx);
    int get_record(double *rec);
    int set_records(const double *recs);
    int solve(const double *x0, double *out_x, eqs_report *rep);

    Ctx *ctx_;
    eqs_pso_params p_{};
    PsoDev d_{};
    std::vector<double> lower_, upper_;
    void *arena_ = nullptr;
    size_t arena_bytes_ = 0;
    double *replay_buf_ = nullptr;
    size_t replay_cap_ = 0;
    PsoCtrl *h_ctrl_ = nullptr; // pinned mirror
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    int grid_step_ = 0, grid_init_ = 0;
    int64_t launches_ = 0;
    int64_t last_launches_ = 0;
    float last_ms_ = 0.f;
    bool manual_ = false;
    bool fused_exchange_ = false; // synchronous step exchanges over peer memory inside the step kernel
    bool in_sync_step_ = false;
    bool initialised_ = false;    // init() / init_apply() has run: stepping or reading before that is refused
    bool seq_windowed_ = false;   // sequential mode: device-driven windowed rounds (one rank, more than PSO_SMALL_MAX particles)
    int grid_seqw_ = 0;
    int max_grid_ = 0;

private:
    int upload_replay(const double *host, int64_t iters);
    int exchange();
    int launch_init_eval();
    int launch_ring_scan();
    int gather_tile_minima();
    int launch_init_apply();
    int launch_step();
    int launch_step_apply();
    int sequential_pass();
    int sequential_windowed(int64_t iters);
    int fetch_ctrl();
};

} // namespace eqs
