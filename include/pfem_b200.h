/* =====================================================================================
 * pfem_b200.h -- C ABI of the B200-native (sm_100a, fp64) PFEM3D finite-element hot path.
 *
 * Drop-in boundary for ImperatorS79/PFEM ("PFEM3D").  The reference has no plugin ABI; its
 * seams are C++ virtuals:
 *     virtual bool Equation::solve()                 srcs/simulation/Equation.hpp:68
 *     MatrixBuilder<dim> get / set vocabulary      srcs/simulation/matricesBuilder/MatricesBuilder.hpp:54-85
 *     the linear solver object m_solver              srcs/simulation/physics/IncompNewton/MomContEquationPSPG.inl:281-290
 * Concrete equations are created only at the REGISTER_EQ sites
 * (physics/IncompNewton/Solver.cpp:6-10,38-40; physics/WCompNewton/Solver.cpp:9-13,45-54).
 * The functions below are what replacement Equation subclasses (shim/pfem_b200_equations.hpp)
 * bind to; each one cites the reference code it replaces.  INTEGRATION.md shows the host patch.
 *
 * Conventions
 *  - plain C, no C++/torch types; every pointer is a HOST pointer owned by the caller and is
 *    not retained after the call returns (device-pointer variants are suffixed _device);
 *  - nodal vectors use the reference's flat layout q[n + s*nNodes]
 *    (getQFromNodesStates, srcs/simulation/utility/StatesFromToQ.hpp:21-34);
 *  - state order: incompressible [u,v,(w),p] (physics/IncompNewton/Problem.cpp:13-14),
 *    weakly compressible [u,v,(w),p,rho,ax,ay,(az)] (physics/WCompNewton/Problem.cpp:17-18);
 *  - global dof of the PSPG system: node + d*nNodes, pressure block last (PSPG.inl:71-86);
 *  - all arithmetic is IEEE fp64; indices are int32 on the device;
 *  - return value: 0 ok; >0 recoverable numerical outcome (the shim maps it to `return false`,
 *    i.e. the reference's dt-halving path, IncompNewton/Solver.cpp:216-224); <0 fatal, message
 *    in pfem_last_error().  Nothing throws across this boundary.
 *  - a context is bound to one GPU, owns one CUDA stream and is not thread-safe; calls are
 *    synchronous on return.  There is NO CPU fallback: without a CUDA device pfem_create fails.
 * ===================================================================================== */
#ifndef PFEM_B200_H
#define PFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFEM_ABI_VERSION 1

/* status codes */
#define PFEM_OK 0
#define PFEM_NOT_CONVERGED 1 /* Krylov / Picard hit maxIter (PicardAlgo.cpp:55-65) */
#define PFEM_NAN 2           /* NaN residual (PicardAlgo.cpp:79-86) or NaN dt (WCompNewton/Solver.cpp:231-232) */
#define PFEM_ERR_INVALID (-1)
#define PFEM_ERR_CUDA (-2)
#define PFEM_ERR_STATE (-3) /* call order: e.g. assemble before set_topology */
#define PFEM_ERR_COMM (-4)

/* node flag bits (Node.hpp:93-105; isFree == node belongs to no element, Node.inl:48-51) */
#define PFEM_NODE_BOUND 1u
#define PFEM_NODE_FREE 2u
#define PFEM_NODE_FIXED 4u
#define PFEM_NODE_FREE_SURFACE 8u

typedef struct pfem_ctx pfem_ctx;

/* material + solver scalars of MomContEqIncompNewton (MomContEquation.inl:28-30, 239-243; dt = Solver::getTimeStep) */
typedef struct pfem_pspg_params {
    double rho, mu, dt;
    double bodyForce[3];
} pfem_pspg_params;

/* continuity-equation variant = the reference's WCompNewton solver id (ContEquation.inl:38-43) */
#define PFEM_WC_CDS_DPDT 0   /* "CDS_dpdt":   dp/dt form, rho from Tait-Murnaghan (ContEquation.inl:353-413)          */
#define PFEM_WC_CDS_DRHODT 1 /* "CDS_drhodt": d(rho)/dt form, p from Tait-Murnaghan (ContEquation.inl:234-301)        */
#define PFEM_WC_CDS_RHO 2    /* "CDS_rho":    mass conservation M(x_new) rho = M(x_old) rho_old (:196-232, 234-301)   */

/* ContEqWCompNewton / MomEqWCompNewton scalars (ContEquation.inl:26-37; MomEquation.inl:28-29, 149-153);
 * meduri != 0  <=>  stabilization == "Meduri" (ContEquation.inl:387-390); eqType: PFEM_WC_CDS_* */
typedef struct pfem_wc_params {
    double mu, K0, K0p, rhoStar;
    double bodyForce[3];
    int32_t meduri;
    int32_t eqType;
} pfem_wc_params;

/* Boussinesq constants (Material.k, cv, alpha, Tr: IncompNewton/MomContEquation.inl:32-38, WCompNewton/HeatEquation.inl:24-27) */
typedef struct pfem_thermal_params {
    double k, cv, alpha, Tr;
} pfem_thermal_params;

typedef struct pfem_info {
    int32_t dim, device, nRanks, rank;
    int64_t nNodes, nElems, nDof;
    int64_t nnzBlocks;     /* (dim+1)x(dim+1) node blocks held on the device          */
    int64_t nnzReference;  /* nnz of the reference's m_A (row masks + explicit zeros)  */
    int64_t deviceBytes;   /* device memory owned by the context                       */
    int32_t maxElemsPerNode, maxNeighbours;
} pfem_info;

/* ---- lifetime ------------------------------------------------------------------- */
int pfem_abi_version(void);
/* dim = Mesh::getDim() (2|3); device = CUDA ordinal.  Fails with PFEM_ERR_CUDA when no device is usable. */
int pfem_create(pfem_ctx** ctx, int dim, int device);
int pfem_destroy(pfem_ctx* ctx);
const char* pfem_last_error(const pfem_ctx* ctx);
/* Run on a caller-provided cudaStream_t (NULL: the context's own stream).  Lets a harness time with its own events. */
int pfem_set_stream(pfem_ctx* ctx, void* cudaStream);
int pfem_get_info(const pfem_ctx* ctx, pfem_info* info);

/* ---- mesh: once per remesh (Mesh::remesh, Mesh.cpp:919-926 stays on the host) ------ */
/* elemNodes: nElems x (dim+1) row-major, as Element::m_nodesIndexes (Element.hpp:121; Element.inl:18-21), positive
 * orientation.  nodeFlags: PFEM_NODE_* per node.  Builds node->element incidence, the node-block sparsity pattern of
 * m_A and (multi-GPU) the partition -- the device counterpart of the pattern work inside setFromTriplets (PSPG.inl:136). */
int pfem_set_topology(pfem_ctx* ctx, int64_t nNodes, int64_t nElems, const uint64_t* elemNodes, const uint8_t* nodeFlags);
/* Node::m_position (Node.hpp:93), layout x[n + d*nNodes] */
int pfem_set_positions(pfem_ctx* ctx, const double* x);
int pfem_get_positions(pfem_ctx* ctx, double* x);
/* Mesh::saveNodesList / restoreNodesList (Mesh.cpp:1019-1053) -- positions only; states are caller-managed */
int pfem_snapshot_positions(pfem_ctx* ctx);
int pfem_restore_positions(pfem_ctx* ctx);
/* Mesh::updateNodesPosition (fromSnapshot=0, Mesh.cpp:1101-1137) / updateNodesPositionFromSave (=1, :1238-1277):
 * x = base + delta for nodes that are not isFixed; delta layout n + d*nNodes. */
int pfem_move_positions(pfem_ctx* ctx, const double* delta, int fromSnapshot);
/* setNodesStatesfromQ / getQFromNodesStates (StatesFromToQ.hpp:9-34): states [first, first+count) */
int pfem_set_states(pfem_ctx* ctx, int first, int count, const double* q);
int pfem_get_states(pfem_ctx* ctx, int first, int count, double* q);
/* Host-evaluated Lua velocity BC: mask[n] != 0 <=> node.isBound() && getBcTagFlags(tag, flag 0); values[n + d*nNodes]
 * = "<type>V"(pos, t+dt) (PSPG.inl:206-214; WCompNewton/MomEquation.inl:355-364). */
int pfem_set_dirichlet(pfem_ctx* ctx, const uint8_t* mask, const double* values);
/* Boundary facets of the current mesh = Mesh::m_facetsList after the alpha shape (Mesh3D.cpp:218-262, Mesh2D.cpp):
 * facetNodes[f*dim + k] = Facet::m_nodesIndexes, outNode[f] = Facet::m_outNodeIndex (the element node in front of the
 * facet), elemIndex[f] = Facet::m_elementIndex.  Call after pfem_set_mesh (a new mesh drops the facets); nFacets = 0
 * clears them.  Facet geometry (Facet::computeJ/DetJ/Normal, Facet.cpp:16-77, 130-210) is recomputed on the device from
 * the current positions whenever it is used. */
int pfem_set_facets(pfem_ctx* ctx, int64_t nFacets, const uint64_t* facetNodes, const uint64_t* outNode,
                    const uint64_t* elemIndex);
/* Surface-tension coefficient gamma (Material.gamma: MomContEquation.inl:30, WCompNewton/MomEquation.inl:29).  With
 * gamma >= 1e-15 and facets set, pfem_pspg_assemble adds MatrixBuilder::getFST (MatricesBuilder.inl:389-403) of every
 * facet with a node on the free surface to the velocity rows of b before the nodal BC pass (PSPG.inl:155-187), and the
 * explicit step adds it to F for facets entirely on the free surface (WCompNewton/MomEquation.inl:312-336).  Default 0
 * (the reference's `if(m_gamma < 1e-15) continue`). */
int pfem_set_surface_tension(pfem_ctx* ctx, double gamma);

/* ---- temperature / shear-rate dependent factors and the heat equations ---- */
/* Problem ids "Boussinesq" / "BoussinesqWC": with thermal parameters set, pfem_pspg_assemble uses the F factor
 * rho (1 - alpha (T - Tr)) and the H factor (1 - alpha (T - Tr)) (MomContEquation.inl:166-199), pfem_wc_step runs
 * m_solveBoussinesqWC (WCompNewton/Solver.cpp:278-320: explicit heat equation HeatEquation.inl:154-298, then continuity,
 * then momentum with the buoyancy factor MomEquation.inl:105-112) and pfem_wc_next_dt takes the thermal diffusivity
 * k/(cv rho) into account (Solver.cpp:214-216).  NULL switches the factors off. */
int pfem_set_thermal(pfem_ctx* ctx, const pfem_thermal_params* t);
/* Problem id "Bingham": K factor mu + tau0 (1 - exp(-mReg gammaDot))/gammaDot (MomContEquation.inl:102-119). */
int pfem_set_bingham(pfem_ctx* ctx, int on, double tau0, double mReg);
/* nodal temperature (the extra node state of the Boussinesq problems), T[nNodes] */
int pfem_set_temperature(pfem_ctx* ctx, const double* T);
int pfem_get_temperature(pfem_ctx* ctx, double* T);
/* mask[n] != 0 <=> getBcTagFlags(node tag, flag 1): a "<type>T" Lua function exists; values[n] = its value
 * (WCompNewton/HeatEquation.inl:203-216, IncompNewton/HeatEquation.inl:383-408) */
int pfem_set_temperature_bc(pfem_ctx* ctx, const uint8_t* mask, const double* values);
/* HeatEqIncompNewton::m_buildAb + m_applyBC (IncompNewton/HeatEquation.inl:227-412, no flux facet terms): scalar system
 * A = M(cv rho) + dt L(k), b = M thetaPrev on the current positions.  Single-GPU contexts. */
int pfem_heat_assemble(pfem_ctx* ctx, double rho, double cv, double k, double dt, const double* thetaPrev);
/* Eigen::ConjugateGradient::solveWithGuess of that system (HeatEquation.inl:129-135): Jacobi-preconditioned CG started
 * from the temperature on the device (zero if none); the solution becomes the device temperature; T may be NULL. */
int pfem_heat_solve(pfem_ctx* ctx, double relTol, int maxIter, double* T, int* iters, double* relRes);
/* the heat system in the reference's format (column-major compressed, masked rows reduced to their diagonal); parity only */
int pfem_heat_export_csc(pfem_ctx* ctx, int64_t* nnz, int32_t* colPtr, int32_t* rowIdx, double* val, double* b);

/* ---- fractional-step solver id "FracStep" (SURVEY 8f rank 1): the three linear systems of one Picard body
 * (MomContEquationFracStep.inl:464-548), each assembled on the device from given inputs, and Eigen's Jacobi-preconditioned
 * conjugate gradients (m_solverIt) for them.  Single-GPU contexts, gamma = 0.
 * 1 velocity prediction: m_buildMatFracStep + m_applyBCVAppStep (:8-298): (M/dt + K) vTilde = F + M/dt v_prev + gammaFS D^T p_prev,
 *   rows of bound / free nodes reduced to the identity, Dirichlet columns eliminated; qPrev = (v_prev, p_prev), (dim+1) nNodes.
 *   Held in the node-block storage of the PSPG system with identity pressure rows: pfem_pspg_export_csc / pfem_pspg_matvec see it.
 * 2 pressure: m_buildMatPcorrStep + m_applyBCPCorrStep (:124-130, 158-166, 300-376): L p = -(rho/dt) D vTilde + gammaFS L p_prev,
 *   rows of free nodes reduced to the identity, zero right-hand side on free-surface nodes; vTilde dim nNodes, pPrev nNodes.
 *   Scalar storage of the heat system: pfem_heat_export_csc sees the matrix, pfem_fs_get_rhs the right-hand side.
 *   (The reference leaves this system singular -- DESIGN.md section 7; it is assembled for parity, pfem_fs_solve on it
 *   returns PFEM_NOT_CONVERGED at the iteration cap like the reference's own solve.)
 * 3 velocity correction: m_buildMatVStep + m_applyBCVStep (:76-83, 167-176, 378-452): M deltaV = dt D^T deltaP, identity rows
 *   for bound / free nodes; the same scalar mass matrix for every component, right-hand side [d][nNodes]. */
int pfem_fs_assemble_vapp(pfem_ctx* ctx, const pfem_pspg_params* p, double gammaFS, const double* qPrev);
int pfem_fs_assemble_pcorr(pfem_ctx* ctx, double rho, double dt, double gammaFS, const double* vTilde, const double* pPrev);
int pfem_fs_assemble_vcorr(pfem_ctx* ctx, double rho, double dt, const double* deltaP);
int pfem_fs_get_rhs(pfem_ctx* ctx, double* b);
/* x = m_solverIt.solve(b) on the system assembled last (:472-478, 491-497, 516-522): zero initial guess, diagonal
 * preconditioner, stop at ||r||^2 < relTol^2 ||b||^2 (Eigen's default relTol is machine epsilon, its iteration cap 2 n);
 * `iters` counts as Eigen's iterations() does.  x: dim nNodes (systems 1, 3: [d][nNodes]) or nNodes (system 2); may be NULL. */
int pfem_fs_solve(pfem_ctx* ctx, double relTol, int maxIter, double* x, int* iters, double* relRes);

/* ---- incompressible PSPG (MomContEqIncompNewton<dim>) ------------------------------- */
/* m_buildAbPSPG + m_applyBCPSPG (PSPG.inl:7-146, 149-235; facet terms: pfem_set_surface_tension).  qPrev: (dim+1)*nNodes. */
int pfem_pspg_assemble(pfem_ctx* ctx, const pfem_pspg_params* p, const double* qPrev);
/* The same split in two, so that a caller that keeps qPrev across Picard iterations (PSPG.inl:273, 298 pass the same
 * qPrevVec[0]) uploads it once: set_qprev copies it to the device, assemble_resident assembles from device-resident data. */
int pfem_pspg_set_qprev(pfem_ctx* ctx, const double* qPrev);
int pfem_pspg_assemble_resident(pfem_ctx* ctx, const pfem_pspg_params* p);
/* Replaces m_solver.analyzePattern/factorize/solve (PSPG.inl:281-290) by a preconditioned Krylov solve on the device:
 * flexible GMRES(40) under the multigrid preconditioner, BiCGSTAB under the Jacobi ones (`iters` counts preconditioner
 * applications for the former, iterations of two applications each for the latter).
 * q (out, (dim+1)*nNodes) may be NULL to leave the solution on the device.  relTol is on ||b - A q|| / ||b||.
 * Returns PFEM_NOT_CONVERGED / PFEM_NAN like a failed factorisation would (PSPG.inl:304-312). */
int pfem_pspg_solve(pfem_ctx* ctx, double relTol, int maxIter, double* q, int* iters, double* relRes);
/* Preconditioner of that solve.  kind: PFEM_PRECOND_AUTO (multigrid when the mesh has more than 32 nodes, else node-block
 * Jacobi), _POINT (diagonal, what Eigen's iterative solvers default to, MomContEquation.hpp:49), _BLOCK (node-block Jacobi),
 * _MG (aggregation multigrid, V(sweeps,sweeps), node-block Jacobi smoothing with the local damping `damping`/r_i, r_i the
 * inf-norm of block row i of D^-1 A).  sweeps <= 0 / damping <= 0 keep the defaults (3 in 3-D, 2 in 2-D; 2.0).  Under _AUTO a multigrid
 * solve that stagnates (less than 1.5 orders of magnitude in 50 iterations) or has not converged after 300 iterations
 * continues with node-block Jacobi.  get_preconditioner reports what the last
 * solve used and its number of multigrid levels (1: none). */
#define PFEM_PRECOND_AUTO 0
#define PFEM_PRECOND_POINT 1
#define PFEM_PRECOND_BLOCK 2
#define PFEM_PRECOND_MG 3
int pfem_pspg_set_preconditioner(pfem_ctx* ctx, int kind, int sweeps, double damping);
int pfem_pspg_get_preconditioner(pfem_ctx* ctx, int* kindUsed, int* levelsOut);
/* ||A q - b||_2 of the currently assembled system (Res::Ax_f, PSPG.inl:368).  q NULL: the last device solution. */
int pfem_pspg_residual(pfem_ctx* ctx, const double* q, double* resAxf);
/* One body of the Picard loop (PSPG.inl:278-313 + :366-370): solve -> node states <- q -> positions = snapshot + dt*v
 * -> reassemble (+BC) -> res = ||A q - b||.  Requires pfem_snapshot_positions + pfem_pspg_assemble first (m_prepare,
 * PSPG.inl:264-277).  q may be NULL. */
int pfem_pspg_picard_iter(pfem_ctx* ctx, const pfem_pspg_params* p, const double* qPrev, double relTol, int maxIter,
                          double* q, double* resAxf, int* iters);
/* The assembled system in the reference's own format: column-major compressed m_A (MomContEquation.hpp:70) with the
 * reference's pattern (row masks, identity rows, explicit zeros of eliminated Dirichlet columns) and m_b.
 * Call with colPtr == NULL to obtain *nnz only.  colPtr: nDof+1. */
int pfem_pspg_export_csc(pfem_ctx* ctx, int64_t* nnz, int32_t* colPtr, int32_t* rowIdx, double* val, double* b);
/* y = A x with the assembled matrix (the sparse*dense product of PSPG.inl:368), host vectors of nDof */
int pfem_pspg_matvec(pfem_ctx* ctx, const double* x, double* y);

/* ---- weakly compressible explicit step --------------------------------------------- */
/* SolverWCompNewton::m_solveWCompNewtonNoT body up to the remesh (WCompNewton/Solver.cpp:236-263): half kick, move,
 * ContEqWCompNewton::solve (CDS_dpdt, ContEquation.inl:123-148), MomEqWCompNewton::solve (MomEquation.inl:201-226).
 * Operates on the device-resident states; nSteps > 1 repeats with the CFL time step recomputed on the device between
 * steps (computeNextDT) when adaptDT != 0. */
int pfem_wc_step(pfem_ctx* ctx, const pfem_wc_params* p, double dt);
/* Kernel formulation of the explicit step (a tuning knob like pfem_pspg_set_preconditioner; results agree to rounding,
 * each variant is bit-reproducible and partition-independent): 0 = chosen by mesh size (default), 6 = one gather kernel
 * per equation (4 lanes per node over the incident elements), 7 = the same with staged neighbour records, 11 = two
 * passes per equation (element records, then an ordered nodal gather; the CFL pass reuses what the step stored),
 * 12 = two-pass continuity + gather momentum, 13 = tiles (a CTA computes the elements of ~64 nearby nodes once and keeps
 * their records in shared memory; CDS_dpdt only). */
int pfem_wc_set_variant(pfem_ctx* ctx, int variant);
/* SolverWCompNewton::computeNextDT (WCompNewton/Solver.cpp:192-234) incl. Element::getRin (Element.cpp:226-294) */
int pfem_wc_next_dt(pfem_ctx* ctx, const pfem_wc_params* p, double securityCoeff, double maxDT, double* dt);
/* nSteps iterations of the Problem::simulate loop body for the explicit solver between two remeshes
 * (Problem.cpp:344-369: solveOneTimeStep then computeNextDT) without returning to the host: *dt is the first time step on
 * entry and the next one on exit, *elapsed the simulated time covered.  One CUDA graph per step on a single-GPU context; on a
 * partitioned mesh the CFL minimum is all-reduced on the device and the steps are enqueued back to back (every rank calls
 * it with the same arguments). */
int pfem_wc_run(pfem_ctx* ctx, const pfem_wc_params* p, int nSteps, double securityCoeff, double maxDT, double* dt,
                double* elapsed);

/* ---- multi-GPU (one context per rank/GPU; elements sharded by RCB) ---- */
/* ncclUniqueId: the 128 opaque bytes of ncclGetUniqueId obtained by rank 0 (pfem_comm_unique_id) and broadcast by the
 * launcher.  Each rank then passes its LOCAL mesh (owned + ghost nodes, one ghost-element layer) to pfem_set_topology
 * followed by pfem_set_partition. */
int pfem_comm_unique_id(void* id128);
int pfem_comm_init(pfem_ctx* ctx, int nRanks, int rank, const void* id128);
/* Halo plan of the LOCAL mesh previously given to pfem_set_topology on this rank (built on the host per remesh, e.g. by
 * pfem_b200/partition.py): nodes [0, nOwned) are owned (their rows are computed here), nodes [nOwned, nNodes) are ghosts
 * grouped by owner.  Peer p receives the owned nodes sendIdx[sendOffsets[p] .. sendOffsets[p+1]) and fills the ghost
 * range [recvStart[p], recvStart[p]+recvCount[p]).  Local elements must be sorted by global element index so that the
 * gather kernels sum in the single-GPU order (sharded results are then bit-identical to one GPU). */
int pfem_set_partition(pfem_ctx* ctx, int64_t nOwned, int nPeers, const int32_t* peerRank, const int64_t* sendOffsets,
                       const int32_t* sendIdx, const int64_t* recvStart, const int64_t* recvCount);
/* In-process communicator: the ranks are contexts of ONE process (the reference is a single process, Problem.cpp:344-369
 * runs its time loop on one main thread), each driven by its own host thread; exchanges are device-to-device copies ordered
 * by CUDA events, the threads meet at barriers inside the collective calls.  Several ranks may share one device (that is
 * how the parity tests run a partitioned mesh on a single GPU).  The group must outlive its contexts.  Every rank must
 * issue the same sequence of API calls. */
int pfem_comm_local_create(int nRanks, void** group);
int pfem_comm_local_destroy(void* group);
int pfem_comm_init_local(pfem_ctx* ctx, void* group, int rank);
/* A rank that failed outside the library releases the others from their barriers (they return PFEM_ERR_COMM). */
int pfem_comm_abort(pfem_ctx* ctx);

/* ---- partitioner (host, C++/OpenMP; once per remesh -- the call site is where the host hands the new mesh over after
 * Mesh::remesh, Mesh.cpp:919-926) ---- */
/* RCB of the node coordinates x[n + d*nNodes] into nRanks parts; a rank keeps every element incident to one of its nodes
 * (one ghost-element layer) in ascending global element index, numbers its owned nodes first (ascending global id) and its
 * ghosts grouped by owner.  elemNodes must stay valid until pfem_partition_destroy (it is read, not copied). */
typedef struct pfem_partition pfem_partition;
int pfem_partition_create(pfem_partition** part, int dim, int64_t nNodes, int64_t nElems, const uint64_t* elemNodes,
                          const double* x, int nRanks);
int pfem_partition_destroy(pfem_partition* part);
int pfem_partition_owner(const pfem_partition* part, int32_t* owner /* nNodes */);
/* sizes of the local mesh of `rank`, then its arrays: l2gNodes[nLocalNodes], l2gElems[nLocalElems], localConn[nLocalElems x
 * (dim+1)] in local node ids, and the halo plan in the layout pfem_set_partition takes (peerRank[nPeers],
 * sendOffsets[nPeers+1], sendIdx[nSend], recvStart[nPeers], recvCount[nPeers]).  Null outputs are skipped. */
int pfem_partition_local_sizes(pfem_partition* part, int rank, int64_t* nLocalNodes, int64_t* nOwned, int64_t* nLocalElems,
                               int32_t* nPeers, int64_t* nSend);
int pfem_partition_local_get(pfem_partition* part, int rank, int64_t* l2gNodes, int64_t* l2gElems, uint64_t* localConn,
                             int32_t* peerRank, int64_t* sendOffsets, int32_t* sendIdx, int64_t* recvStart, int64_t* recvCount);

/* ---- instrumentation (phase names = the reference's m_accumalatedTimes keys, PSPG.inl:19-369) ---- */
/* on: 0 off | 1 phases | 2 phases + per-kernel phases of the multigrid cycle (which then runs un-graphed) */
int pfem_profile_enable(pfem_ctx* ctx, int on);
int pfem_profile_reset(pfem_ctx* ctx);
/* accumulated device milliseconds (CUDA events on the context's stream) and call count of one phase */
int pfem_profile_get(pfem_ctx* ctx, const char* phase, double* ms, int64_t* calls);
/* number of kernels this library has launched on the context since creation */
int pfem_launch_count(const pfem_ctx* ctx, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* PFEM_B200_H */
