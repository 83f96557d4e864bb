/*
 * jme_b200.h - C ABI of libjme_b200.so: the B200 (sm_100a) drop-in for JMolecularEnergy's
 * MMFF94 energy / gradient hot path.
 *
 * Plain C, no CUDA or torch types: every pointer is caller-owned HOST memory unless the
 * name says "_device".  One handle = one parameterised molecule on one GPU (the current
 * CUDA device at jme_create).  A handle is not thread-safe, like the reference's
 * ForceField objects.  All functions return JME_OK (0) or a negative jme_status; the text
 * of the last failure is in jme_last_error(handle) (jme_last_create_error() when no
 * handle exists yet).  There is no CPU fallback: without a usable CUDA device jme_create
 * fails with JME_ERR_CUDA.
 *
 * Reference interfaces replaced (paths under /root/reference/src/main/java/org/jme):
 *   jme_create            <- the result of MMFF94.assignParameters, forcefield/mmff/MMFF94.java:111-219
 *                            (atom/bond/container properties set at :117-120,:151,:160,:215-219)
 *   jme_set_options       <- ForceField.setCutoffDistance/setDielectricConstant/setEnergyUnit,
 *                            forcefield/ForceField.java:175,199,117
 *   jme_set_coords        <- reading IAtom.getPoint3d() of every atom (e.g. MMFF94VdwComponent.java:136)
 *   jme_set_coord1        <- GradientDescentMinimizer.movePoint, minimizer/GradientDescentMinimizer.java:107-115
 *   jme_energy            <- MMFF94.calculateEnergy, forcefield/mmff/MMFF94.java:385-397 and the seven
 *                            EnergyComponent.calculateEnergy overrides it sums (forcefield/EnergyComponent.java:55)
 *   jme_energy_gradient   <- no counterpart: the reference has no gradient; this is
 *                            d(jme_energy)/d(xyz)
 *   jme_last_error        <- the unchecked exceptions of MMFF94.java:371-375 etc.
 */
#ifndef JME_B200_H
#define JME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JME_ABI_VERSION 1

typedef struct jme_handle jme_handle_t;

typedef enum {
    JME_OK = 0,
    JME_ERR_INVALID = -1,      /* bad argument (null pointer, index out of range, size mismatch) */
    JME_ERR_CUDA = -2,         /* no device / CUDA runtime failure (text has the CUDA error) */
    JME_ERR_NOT_ASSIGNED = -3, /* coordinates never set: "MMFF94 parameters need to be assigned first" */
    JME_ERR_NOMEM = -4,
    JME_ERR_UNSUPPORTED = -5
} jme_status;

/* ForceField.EnergyUnit (ForceField.java:133-160); factors 4184 / 1000 / 1 J */
typedef enum { JME_KCAL_PER_MOL = 0, JME_KJ_PER_MOL = 1, JME_JOULE_PER_MOL = 2 } jme_energy_unit;

/* Component order of MMFF94.java:98 */
enum { JME_BOND = 0, JME_ANGLE = 1, JME_STRETCH_BEND = 2, JME_OOP = 3, JME_TORSION = 4, JME_VDW = 5, JME_ELE = 6,
       JME_N_COMPONENTS = 7 };

/* Nonbonded path selection (jme_set_path). AUTO: all-pairs tiles when cutoff <= 0 or the
 * system is small, cell-list path otherwise.  The cell-list path has two evaluation schemes, chosen by
 * size under JME_PATH_CELLS: per-atom survivor lists (small systems) and cluster lists with Newton's third
 * law (large systems); the last two values force one of them (tests, tuning). */
typedef enum { JME_PATH_AUTO = 0, JME_PATH_ALLPAIRS = 1, JME_PATH_CELLS = 2, JME_PATH_CELLS_LISTS = 3,
               JME_PATH_CELLS_CLUSTERS = 4 } jme_path;

/*
 * Everything MMFF94.assignParameters leaves on the container, flattened.  All tabulated
 * parameters are binary32 because the reference stores them as float/Float
 * (MMFF94Parameters.java:361-379,432-444) and widens at use; pass the floats, never
 * re-parsed decimals.
 */
typedef struct {
    int64_t n_atoms;
    const int32_t* atom_type;   /* [n_atoms] MMFF numeric type 1..82, 87..99 ("mmff.type", MMFF94.java:117) */
    const double*  charge;      /* [n_atoms] partial charge, atom.getCharge() (MMFF94.java:616) */
    const uint8_t* linear;      /* [n_atoms] GeometricParameters.ideallyLinear of the atom (MMFF94.java:120) */

    int32_t n_vdw_types;        /* rows of MMFF94-vdw.par for the types that occur */
    const int32_t* vdw_type;    /* [n_vdw_types] numeric type of the row */
    const float* vdw_alpha;     /* VdwParameters.alpha, N, A, G (MMFF94Parameters.java:368) */
    const float* vdw_N;
    const float* vdw_A;
    const float* vdw_G;
    const uint8_t* vdw_DA;      /* 0 none, 1 donor, 2 acceptor (MMFF94Parameters.java:363-365) */

    int64_t n_bonds;            /* container.bonds() with StretchParameters (MMFF94.java:222-256) */
    const int32_t* bond_i; const int32_t* bond_j;
    const float* bond_kb; const float* bond_r0;

    int64_t n_angles;           /* "mmff.angleBending" + "mmff.stretchBend" maps share keys (MMFF94.java:180-183) */
    const int32_t* angle_i; const int32_t* angle_j; const int32_t* angle_k;   /* j is the centre */
    const float* angle_ka; const float* angle_theta0;                         /* Float[4], Float[5] of the angle row */
    const float* stbn_kba_ijk; const float* stbn_kba_kji;                     /* Float[4], Float[5] of the stretch-bend row */

    int64_t n_oops;             /* non-null rows of "mmff.outOfPlane" (MMFF94.java:184-193) */
    const int32_t* oop_i; const int32_t* oop_j; const int32_t* oop_k; const int32_t* oop_l;  /* j centre, l out of plane */
    const float* oop_koop;

    int64_t n_torsions;         /* "mmff.torsion" (MMFF94.java:196-210) */
    const int32_t* tor_i; const int32_t* tor_j; const int32_t* tor_k; const int32_t* tor_l;
    const float* tor_v1; const float* tor_v2; const float* tor_v3;
} jme_system_t;

int jme_abi_version(void);
const char* jme_last_create_error(void);

/* Marshal a parameterised molecule to the current CUDA device: SoA coordinates, charges, dense
 * type indices, type-pair vdW table, bonded term arrays, and the nonbonded pair predicate
 * (1-2/1-3 excluded, 1-4 flagged; MMFF94.java:162-170,1106-1140) rebuilt from the bond list. */
int jme_create(const jme_system_t* sys, jme_handle_t** out);
void jme_destroy(jme_handle_t* h);
const char* jme_last_error(const jme_handle_t* h);

/* cutoff <= 0 means unlimited, dielectric > 0 (ForceField.java:162-210). */
int jme_set_options(jme_handle_t* h, double cutoff, double dielectric, int energy_unit);
int jme_set_path(jme_handle_t* h, int path);
/* The scheme the last evaluation used: JME_PATH_ALLPAIRS, JME_PATH_CELLS_LISTS or JME_PATH_CELLS_CLUSTERS
 * (JME_PATH_AUTO before the first evaluation). */
int jme_active_path(const jme_handle_t* h);

/* Atom-row sharding for multi-GPU runs: this handle evaluates rank `rank` of `world` shares of the
 * nonbonded rows and bonded terms; the caller sums energies and gradients over ranks (all-reduce). */
int jme_set_shard(jme_handle_t* h, int rank, int world);

/* xyz is AoS [n_atoms][3] (x0,y0,z0,x1,...) like an array of Point3d. */
int jme_set_coords(jme_handle_t* h, const double* xyz, int64_t n_atoms);
int jme_set_coord1(jme_handle_t* h, int64_t atom, int axis, double value);
int jme_get_coords(jme_handle_t* h, double* xyz, int64_t n_atoms);

/* total = sum of comps[0..6] in component order, each already in the selected energy unit. */
int jme_energy(jme_handle_t* h, double* total, double comps[JME_N_COMPONENTS]);
/* grad_xyz AoS [n_atoms][3] = dE/dx (NOT the force), same unit per Angstrom. */
int jme_energy_gradient(jme_handle_t* h, double* total, double comps[JME_N_COMPONENTS], double* grad_xyz);

/* pairs_evaluated: non-excluded pairs inside the cutoff whose vdW+ele terms were summed in the last
 * evaluation of this shard; pairs_candidate: non-excluded pairs before the cutoff test (all-pairs
 * path) or distance tests made (cell path). */
int jme_pair_stats(jme_handle_t* h, uint64_t* pairs_evaluated, uint64_t* pairs_candidate);

/* Test/diagnostic: the selected pair set of the last evaluation, i<j, is14 flag; returns the number
 * of pairs through *n_pairs; fills at most `capacity` entries. */
int jme_pair_list(jme_handle_t* h, int32_t* pair_i, int32_t* pair_j, uint8_t* is14, int64_t capacity,
                  int64_t* n_pairs);

/* ---- device-resident entry points (multi-GPU plumbing and HBM-resident benchmarking) ----
 * All work is enqueued on `stream`: a cudaStream_t passed as void* - NULL is the legacy default stream,
 * JME_STREAM_OWN selects the handle's own non-blocking stream again (the state after jme_create);
 * nothing is synchronised.  d_xyz: device AoS [n_atoms][3].  d_out: device doubles
 * [0]=total [1..7]=components [8]=pairs evaluated [9]=pairs candidate (both as double)
 * [10..10+3n) gradient AoS (only when want_gradient).  This is the buffer a multi-GPU caller all-reduces.
 * Nothing can be returned from an asynchronous call, so a failure of the cutoff path that only the device
 * sees - a cutoff sphere with more atoms than the survivor-list capacity - sets the total ([0]) to NaN; the
 * host entry points (jme_energy, jme_energy_gradient, jme_pair_list) grow the lists and re-evaluate instead,
 * after which the device entry points work at the new capacity too. */
#define JME_STREAM_OWN ((void*)(intptr_t)-1)
int jme_set_stream(jme_handle_t* h, void* stream);
int jme_set_coords_device(jme_handle_t* h, const double* d_xyz, int64_t n_atoms);
int jme_eval_device(jme_handle_t* h, int want_gradient, double* d_out);
int64_t jme_eval_out_size(const jme_handle_t* h, int want_gradient);   /* doubles */

/* The reference's coordinate-wise minimizer loop, restated natively so that BASELINE config 4 (energy calls per
 * second inside the optimizer) can be measured without a JVM: GradientDescentMinimizer.minimize
 * (minimizer/GradientDescentMinimizer.java:49-105) with no constraint - per coordinate: move, full energy,
 * keep / undo, slope *= 1.05 / 0.9; one initialising sweep, then `max_steps` sweeps or |dE| <= threshold.
 * Every trial is one jme_set_coord1 + one energy evaluation on the device.  xyz_inout: AoS [n][3], updated in
 * place; step_energies[k] = energy after step k (k = 0 is the initial energy), at most `capacity` entries. */
int jme_gd_minimize(jme_handle_t* h, int max_steps, double step_size, double threshold, double* xyz_inout,
                    int64_t* energy_calls, int* steps_done, double* step_energies, int capacity);

/* AdamMinimizer.minimize (minimizer/AdamMinimizer.java:62-120, the AdamOptimizer of :132-161) restated the same
 * way, no constraint: per coordinate one energy before the move, the Adam step, one energy after it, keep / undo.
 * A step that rounds to zero gives an infinite or NaN slope exactly as the Java division does. */
int jme_adam_minimize(jme_handle_t* h, int max_steps, double step_size, double threshold, double beta1, double beta2,
                      double epsilon, double* xyz_inout, int64_t* energy_calls, int* steps_done, double* step_energies,
                      int capacity);

/* The energy CHANGE of a single-coordinate move - what the reference's minimizers compute as the difference of two
 * full evaluations (GradientDescentMinimizer.java:75-79, AdamMinimizer.java:89-93) - from the terms that contain the
 * moved atom only: its N - 1 nonbonded pairs (same exclusions, 1-4 scaling, cutoff predicate) and its bonded terms,
 * each at the old and the new position, with the arithmetic of the full evaluation.  O(N).  Performs the move (like
 * jme_set_coord1) and returns delta[0] = change of the total, delta[1..7] = change of the seven components, in the
 * handle's energy unit.  Addition - the reference has nothing like it.
 * jme_set_incremental(h, 1) makes jme_gd_minimize / jme_adam_minimize use it for every trial move (one full evaluation
 * at the start only); mode 2 runs the whole loop in one kernel - a cluster of eight thread blocks that keep the
 * coordinates in shared memory and exchange their partial sums through distributed shared memory, no host round trip
 * per trial - for systems of up to 8 192 atoms (larger ones run as mode 1); the default (0) evaluates the full energy
 * per trial like the reference. */
int jme_energy_delta(jme_handle_t* h, int64_t atom, int axis, double value, double* delta /*[8]*/);
int jme_set_incremental(jme_handle_t* h, int mode);

/* ---- owner-mode exchange of a multi-GPU job over NVLink peer memory ----
 * One process per GPU, every rank owns chunk = jme_peer_chunk() consecutive atoms (by original index; the last
 * ranks' chunks are padded).  Each rank creates a communicator on its current device, hands the 64-byte handle of
 * its exported block (a cudaIpcMemHandle_t) to the others by whatever means the host has (MPI, torch.distributed,
 * a JVM-side socket), and opens theirs (`handles`: world x 64 bytes in rank order; the own entry is ignored).
 * jme_peer_step then replaces all-gather + all-reduce + reduce-scatter around the local evaluation of `h`
 * (jme_set_shard(h, rank, world) already applied): d_xyz_own = device AoS [chunk][3] of this rank's atoms,
 * d_grad_own receives [chunk][3], d_header the 10 job-wide doubles of jme_eval_device's header; everything is
 * enqueued on h's stream, nothing is synchronised.  Every rank must call it the same number of times.  The sum
 * over ranks runs in rank order (reproducible).  A rank that waits more than 4 s for a peer gives up and the total
 * ([0] of the header) is NaN. */
typedef struct jme_peer_comm jme_peer_comm_t;
int jme_peer_create(int rank, int world, int64_t n_atoms, jme_peer_comm_t** out);
int64_t jme_peer_chunk(const jme_peer_comm_t* c);
int jme_peer_handle(const jme_peer_comm_t* c, void* handle64);
int jme_peer_open(jme_peer_comm_t* c, const void* handles);
int jme_peer_step(jme_peer_comm_t* c, jme_handle_t* h, const double* d_xyz_own, double* d_grad_own, double* d_header);
const char* jme_peer_last_error(const jme_peer_comm_t* c);
void jme_peer_destroy(jme_peer_comm_t* c);

/* Kernel launches issued by this handle since creation (bench.py's gpu_launches). */
uint64_t jme_launch_count(const jme_handle_t* h);

/* Kernel timing for the roofline: when on, every evaluation records CUDA events (on the launching
 * stream) before the cell build, before and after the nonbonded pair kernel.  jme_profile_read (call it
 * after the stream is synchronised) returns the per-evaluation durations recorded since the last read:
 * nonbonded_ms[k] = pair kernel (all-pairs tiles, or the FP64 evaluation of the survivor lists),
 * build_ms[k] = cell sort + survivor-list build of the cutoff path (0 on the all-pairs path). */
int jme_set_profiling(jme_handle_t* h, int on);
int jme_profile_read(jme_handle_t* h, float* nonbonded_ms, float* build_ms, int capacity, int* n_read);

/* FP64 FMA-pipe peak of the current device measured with a dependent-chain DFMA kernel:
 * returns TFLOP/s (2 flop per FMA) through *tflops. */
int jme_measure_fp64_peak(double* tflops, double* sm_clock_mhz_estimate);

/* Host-side topology helpers (marshalling; MMFF94.java:158-212,1106-1140).  Two-call pattern:
 * pass NULL outputs to get counts. */
int jme_enumerate_terms(int64_t n_atoms, const int32_t* atom_type, const uint8_t* linear,
                        int64_t n_bonds, const int32_t* bond_i, const int32_t* bond_j,
                        int32_t* angles /*[n][3]*/, int64_t* n_angles,
                        int32_t* oops /*[n][4]*/, int64_t* n_oops,
                        int32_t* torsions /*[n][4]*/, int64_t* n_torsions);
int jme_exclusions(int64_t n_atoms, int64_t n_bonds, const int32_t* bond_i, const int32_t* bond_j,
                   int64_t* row_ptr /*[n_atoms+1]*/, int32_t* partner, uint8_t* is14, int64_t capacity,
                   int64_t* n_entries);

#ifdef __cplusplus
}
#endif
#endif /* JME_B200_H */
