// Preconditioner handles (Jacobi, ILU(0)) and the level-scheduled triangular solve plans.
#pragma once
#include "matrix.cuh"

struct phb200_trsv_plan {
    int lower;
    int64_t n;
    int64_t n_levels;
    int32_t* order;        // device: rows sorted by level
    int32_t* level_ptr;    // device: n_levels+1 offsets into order
    int32_t* h_level_ptr;  // host copy
    int32_t max_level_rows;
    // sync-free execution: every level padded to a multiple of 32 rows (-1 = no row), so a warp never spans two levels
    int32_t* order_pad;    // device
    int64_t npad;
    int32_t* warp_level;   // device: level of every 32-row chunk of order_pad
    int32_t* level_warps;  // device: number of 32-row chunks per level
    unsigned long long* level_cnt;   // device: chunks of the level finished so far, accumulated over all solves (epoch * level_warps when complete)
    int epoch;             // host-side counter, bumped per kernel
};

struct phb200_precond {
    int kind;              // PHB200_PRE_*
    int dtype;
    int64_t n;
    const void* inv_diag;  // Jacobi (caller-owned)
    int is_const;          // Jacobi: every entry of inv_diag equals const_inv (constant-diagonal stencils)
    double const_inv;
    // ILU(0): factors share A's pattern; values in lu_val (caller-owned); diag_pos[i] = index of (i,i) in row i
    phb::CsrView lu;
    int32_t* diag_pos;     // device, owned
    phb200_trsv_plan* plan_l;
    phb200_trsv_plan* plan_u;
    // star-stencil wavefront solves (star_trsv.cuh): reciprocal pivots + block counters; star_mode 0 = CSR level solves,
    // 1 = skewed warps, 2 = row scans with CTA barriers, 3 = row scans as a warp dataflow, 4 = x-marching row pipelines (default in 3-D)
    int star_mode;
    int32_t* sc_order;                 // device, owned: tiles of the row-scan kernel in dependency (anti-diagonal) order
    int32_t* pp_order;                 // device, owned: (row tile, z chunk) CTAs of the x-marching pipelines in dependency order
    void* pp_buf;                      // device, owned: hand-over words of the pipelines (row above a tile, carry into a chunk)
    int64_t pp_words;
    unsigned pp_epoch;                 // tags handed out so far
    phb::StarDesc star;
    void* rd;                          // device, owned: 1 / U_ii
    void* rd_l;                        // device: reciprocal pivots that define L (L_ik = A_ik rd_l[k]); == rd for the exact factors,
                                       // owned and distinct for fixed-point sweeps (there L lags the diagonal by one sweep)
    int sweeps;                        // -1: exact ILU(0); k >= 1: k fixed-point sweeps of incomplete_lu_coo
    unsigned long long* tr_flags;      // device, owned
    int64_t tr_flag_cap;               // counters allocated
    unsigned* tr_ticket;               // device, owned
    unsigned long long tr_epoch;       // host: value all counters of a finished launch stand at
    // explicit coarse ('cluster') preconditioner, phiml/backend/_linalg.py:734-766 — all arrays caller-owned
    const int32_t* cl_of;              // cluster of every element
    const int32_t* cl_members;         // element indices grouped by cluster
    const int32_t* cl_ptr;             // cl_count + 1 offsets into cl_members
    const void* cl_inv;                // cl_count x cl_count inverse coarse matrix, row-major
    const void* cl_size;               // cluster sizes in the vector dtype
    int cl_count;
    void* cl_scratch;                  // owned: 2 x cl_batch_cap x cl_count values (coarse vector, coarse solution)
    int64_t cl_batch_cap;
};

namespace phb {

// out = M^-1 in for (batch, ld) vectors; tmp is a scratch vector set of the same shape (ILU only).
int build_plan_star(phb200_trsv_plan** out, const StarDesc& st, int lower, cudaStream_t stream);
int precond_apply(const phb200_precond* p, const void* in, void* out, void* tmp, int64_t batch, int64_t ld, cudaStream_t stream);
// out = L^-1 in (ILU) / D^-1 in (Jacobi) — Preconditioner.apply_inv_l
int precond_apply_inv_l(const phb200_precond* p, const void* in, void* out, int64_t batch, int64_t ld, cudaStream_t stream);

}  // namespace phb
