/* ek_adjoint.h -- job/term descriptors of the level-batched adjoint kernel (ek_adjoint.cu) */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

#define EK_ADJ_CHUNK 2048u     /* elements of one (job, chunk) work item */

/* operand kinds (2 bits each in EkAdjTerm::flags: weight = bits 0-1, adjoint = bits 2-3) */
#define EK_ADJ_ARRAY  0u       /* device array of the job's size        */
#define EK_ADJ_SCALAR 1u       /* device pointer to one value           */
#define EK_ADJ_IMM    2u       /* value bits stored in the descriptor   */

struct EkAdjTerm {             /* one out-edge: contribution weight * adjoint(target) */
    uint64_t w;                /* weight: pointer or immediate bits                    */
    uint64_t g;                /* adjoint of the edge's target: pointer or immediate   */
    uint32_t flags;
    uint32_t pad;
};

struct EkAdjJob {              /* one source node */
    uint64_t dst;              /* adjoint of the source (written once)                 */
    uint32_t first_term, n_terms;
    uint32_t size;             /* elements                                             */
    uint32_t aligned;          /* all array operands 16-byte aligned                   */
};

cudaError_t ek_launch_adjoint(bool f64, const EkAdjJob *jobs, const EkAdjTerm *terms,
                              const uint32_t *chunk_start, uint32_t n_jobs, uint32_t n_chunks,
                              unsigned grid, cudaStream_t stream);
