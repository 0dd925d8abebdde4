// Shell-level host machinery of libqbx.so: shell reconstruction from the flat primitive
// table, shell-pair data, Schwarz bounds, per-class shell-quartet task lists, the packed
// ERI store and the Fock-build driver.  Device kernels live in eri_class.cuh / digest.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "eri_class.cuh"

struct HostShell {
    int l = 0;
    double cen[3] = {0, 0, 0};
    std::vector<double> xpn, coef;       // radial contraction (component-independent)
    int bf[6] = {-1, -1, -1, -1, -1, -1}; // basis-function index per Cartesian component
    double scale[6] = {1, 1, 1, 1, 1, 1}; // per-component weight factor
};

// arguments of the per-class consumers of a packed value block
struct DigestArgs {
    const int4 *bra_info, *ket_info;  // per pair: (shell A, shell B, first internal function of A, of B)
    const int2 *tasks;
    int64_t ntasks;
    const double *vals;          // [ncomp][ntasks]
    int nbf, nmat, same_class;   // nbf = internal dimension here
    int spread;                  // number of block slots the task list is dealt over (see digest.cuh)
    const double *DJ, *DK;       // internal numbering: nbf^2, nmat * nbf^2
    double *Jt, *Kt;             // nbf^2, nmat * nbf^2 (half-accumulators, see digest.cuh)
};
struct ScatterArgs {
    const int2 *bra_shells, *ket_shells;
    const int2 *tasks;
    int64_t ntasks;
    const double *vals;
    const int *shell_bf;
    int64_t nbf;
    double *tensor;              // nbf^4, column-major
};

struct ClassOps {
    int la, lb, lc, ld, ncomp;
    int (*eri)(const ClassArgs &, cudaStream_t);
    int (*digest)(const DigestArgs &, cudaStream_t);
    int (*scatter)(const ScatterArgs &, cudaStream_t);
    int (*eri_split)(const ClassArgs &, cudaStream_t);     // one warp per task (diagonal classes only, else null)
};
const ClassOps *qbx_class_ops(int bra_cls, int ket_cls);    // pair class = la (la + 1) / 2 + lb

struct DevPairSet {
    int la = 0, lb = 0, npair = 0, nprim = 0;
    int2 *shells = nullptr;
    int *prim_off = nullptr;
    double *geom = nullptr, *prim = nullptr, *schwarz = nullptr, *soa = nullptr;
    int2 *soa_idx = nullptr;
    int4 *info = nullptr;                // (shell A, shell B, first internal function of A, of B)
    std::vector<int> h_nprim;            // primitive pairs per pair (host copy, for cost models)
    PairSet view() const { return PairSet{shells, prim_off, geom, prim, soa, soa_idx, npair}; }
};

struct TaskList {
    int2 *tasks = nullptr;               // (bra pair, ket pair); ket = -1 marks an unused slot of a group task
    int64_t n = 0;                       // slots
    int64_t nvalid = 0;                  // real quartets among them
    double nprimq = 0;                   // primitive quartets behind these tasks
    // group tasks (eri_group.cu): task t covers the slots gt_off[t] .. gt_off[t] + nmem(gt_grp[t]) - 1
    int ngt = 0;
    int *gt_bra = nullptr, *gt_grp = nullptr, *gt_off = nullptr;
    int *order = nullptr;                // 32-task chunks sorted by cost, heaviest first (LPT schedule for the work queue)
    int nheavy = 0;                      // leading chunks of `order` whose tasks exceed QBX_HEAVY_TASK primitive quartets
};
// device scratch of one class between the count and the fill phase of task building
struct TaskScratch {
    int *cnt[3] = {nullptr, nullptr, nullptr};
    int64_t *off[3] = {nullptr, nullptr, nullptr};
};
// cost-sorted chunk order of a task list (tasks != null) or of a group-task list (tasks == null);
// enqueue-only, *d_nheavy (device) += chunks above QBX_HEAVY_TASK
int qbx_chunk_order(const int2 *tasks, const int *gt_bra, const int *gt_grp, int64_t n, const int *poff_bra,
                    const int *poff_ket, int **order_out, int *d_nheavy, cudaStream_t s);
// off[i] = exclusive prefix sum of cnt, off[n] = *d_total = the sum (enqueue-only)
void qbx_scan_counts(const int *cnt, int n, int64_t *off, int64_t *d_total, cudaStream_t s);
#define QBX_HEAVY_TASK 1024.0f

// (ss) group pairs: ket-side general-contraction sharing, see eri_group.cu
struct GroupSet {
    int ng = 0;
    int *nmem = nullptr, *members = nullptr, *prim_off = nullptr;
    int *flip = nullptr;                 // bit m: member m lists its shells as (shell of Q, shell of P) -- the pair list orders by shell index
    double *soa = nullptr;
    int2 *soa_idx = nullptr;
    std::vector<int> h_nprim, h_nmem;
};
int qbx_group_build(const std::vector<HostShell> &sh, const std::vector<int2> &ss_pairs, GroupSet &out,
                    const int2 *d_ss_pairs);             // primitive records computed on the device from the (ss) pair list
void qbx_group_free(GroupSet &g);
// two enqueue-only phases, see Engine::tasks_count / tasks_fill.  d_total[0..2] = group tasks of all
// ranks, group tasks of this rank, slots of this rank.
int qbx_group_count(const GroupSet &G, const struct DevPairSet &B, const struct DevPairSet &K, bool same, double tol, int rank,
                    int nranks, TaskScratch &ts, int64_t *d_total, cudaStream_t s);
int qbx_group_fill(const GroupSet &G, const struct DevPairSet &B, const struct DevPairSet &K, bool same, double tol, int rank,
                   int nranks, TaskScratch &ts, const int64_t *h_total, TaskList &tl, double *d_stat, int *d_nheavy,
                   cudaStream_t s);
int qbx_group_eri(int la, const GroupSet &G, const ClassArgs &a, const TaskList &tl, cudaStream_t s);
int qbx_group_digest(int la, const GroupSet &G, const DigestArgs &a, const TaskList &tl, cudaStream_t s);   // -1: not served

class Engine {
public:
    static Engine *create(int64_t nprim, const double *cen, const double *xpn, const int32_t *ang, int64_t nbf,
                          const int64_t *bf_off, const int64_t *bf_prim, const double *bf_w);
    static Engine *from_shells(const std::vector<HostShell> &shells, int64_t nbf, bool pair_adjacent);
    ~Engine();

    void info(int64_t *info) const;
    int fill_tensor(double *d_tensor, cudaStream_t s, double *stats);
    int store(double tol, int mode, int rank, int nranks, cudaStream_t s, double *stats);
    int recompute(cudaStream_t s, double *stats, bool timed = false);   // enqueue only; timed = serialised, per-class events
    int class_stats(cudaStream_t s, double *stats, double *out);        // runs one serialised, timed recompute
    int fock(int nmat, const double *dDJ, const double *dDK, double *dG, cudaStream_t s, double *stats);
    void release_store();

    static int synthetic(int la, int lb, int lc, int ld, int K, int64_t nq, uint64_t seed, double *secs,
                         double *checksum, double *prim_quartets, int64_t nsample, double *sample_out,
                         double *sample_geom, cudaStream_t s);

private:
    Engine() = default;
    int upload(bool pair_adjacent);
    int ensure_schwarz(cudaStream_t s);
    int build_tasks(int bc, int kc, double tol, int rank, int nranks, TaskList &out, cudaStream_t s);
    int tasks_count(int bc, int kc, double tol, int rank, int nranks, bool grp, TaskScratch &ts, int64_t *d_total, cudaStream_t s);
    int tasks_fill(int bc, int kc, double tol, int rank, int nranks, bool grp, bool want_order, TaskScratch &ts,
                   const int64_t *h_total, TaskList &out, double *d_stat, int *d_nheavy, cudaStream_t s);
    int run_eri(int bc, int kc, const int2 *tasks, int64_t n, double *out, cudaStream_t s, const int *order = nullptr);
    int eri_args(int bc, int kc, const int2 *tasks, int64_t n, double *out, cudaStream_t s, ClassArgs &a);
    // (ss|ss), (ps|ss); (ds|ss) through the group kernel was measured slower twice (54 accumulators per thread:
    // 41.4 vs 40.8 ms per step in round 2) and is not routed there
    bool grouped(int bc, int kc) const { return use_groups_ && kc == 0 && (bc == 0 || bc == 1); }
    GroupSet groups_;
    bool use_groups_ = false;

    std::vector<HostShell> shells_;
    int64_t nbf_ = 0;
    int maxl_ = 0;
    int *d_shell_bf_ = nullptr;
    double *d_shell_scale_ = nullptr;
    DevPairSet pairs_[QBX_NPAIRCLS];
    bool have_schwarz_ = false;
    // stored state
    int mode_ = -1;
    TaskList tasks_[QBX_NPAIRCLS][QBX_NPAIRCLS];
    double *vals_[QBX_NPAIRCLS][QBX_NPAIRCLS] = {{nullptr}};
    double *chunk_ = nullptr;            // direct mode / tensor fill staging
    int64_t chunk_doubles_ = 0;
    double *d_Jt_ = nullptr, *d_Kt_ = nullptr;
    int *d_bad_ = nullptr;               // set by k_permute_in when a density is not symmetric
    int *d_shell_first_ = nullptr, *d_ext_of_int_ = nullptr;
    int64_t nint_ = 0;                   // internal dimension (complete Cartesian shells)
    double *d_Dint_ = nullptr;           // DJ, DK[0], DK[1] in internal numbering
    int64_t n_quartets_ = 0, n_values_ = 0, stored_bytes_ = 0;
    double n_primq_ = 0, model_flops_ = 0;
    // the 21 class kernels of a pass are independent: they are dealt over a few side streams so
    // that the tail of one class overlaps the head of the next
#ifndef QBX_SIDE_STREAMS
#define QBX_SIDE_STREAMS 4
#endif
    static constexpr int kSide = QBX_SIDE_STREAMS;
    cudaStream_t side_[kSide] = {nullptr};
    cudaEvent_t side_ev_[kSide] = {nullptr}, fork_ev_ = nullptr;
    int fork(cudaStream_t s);
    // stored-mode Fock build as a CUDA graph: the ~30 launches (permute, clears, 21 digestions over the side streams,
    // finish) of a build are captured the SECOND time the same (nmat, DJ, DK, G) device pointers are seen on the same
    // store and replayed from then on -- an SCF calls the build 30 times with fixed buffers.  QBX_FOCK_GRAPH=0: eager.
    cudaGraphExec_t fock_graph_ = nullptr;
    struct FockKey { int nmat = 0; const void *dj = nullptr, *dk = nullptr; void *g = nullptr; cudaStream_t s = nullptr; int64_t gen = -1; } fock_key_;
    int fock_seen_ = 0;                  // consecutive eager builds with fock_key_
    double fock_stats_[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // what one captured build adds to the handle's counters
    int64_t store_gen_ = 0;              // bumped whenever the lists / values are rebuilt or released
    int fock_enqueue(int nmat, const double *dDJ, const double *dDK, double *dG, cudaStream_t s, double *stats);
    void drop_fock_graph();
    int join(cudaStream_t s);
    unsigned int *d_counters_ = nullptr; // work-queue heads, one per ERI launch (rotating)
    int counter_next_ = 0;
    cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
    cudaEvent_t cls_ev_[QBX_NCLASS + 1] = {nullptr};
    bool cls_timed_ = false;
};

int qbx_launch_eri_coop(int la, int lb, int lc, int ld, const ClassArgs &a, cudaStream_t s);   // eri_coop.cu
double qbx_model_flops_prim(int la, int lb, int lc, int ld);   // prim + acc of SURVEY.md 8(d)
double qbx_model_flops_hrr(int la, int lb, int lc, int ld);
