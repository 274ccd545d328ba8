/*
 * pqp_oracle_arena.h -- CPU ORACLE (test infrastructure, not product code; see pqp_oracle.h).
 *
 * Per-thread bump arena for the scratch of one oracle solve.  The batch driver (pqp_oracle_api.c), which is
 * what bench.py times as the CPU baseline, switches it on around every path: the ~80 malloc/calloc/free
 * calls of an assembly + OSQP solve then cost a pointer bump each instead of going through the C library's
 * (lock-contended, at 64-128 threads) allocator, and the arena is simply reset between paths.  With the
 * arena off (single solves from the tests, the shared symbolic cache) the calls go to the C library as
 * before.  Nothing numerical depends on it.
 */
#ifndef PQP_ORACLE_ARENA_H_
#define PQP_ORACLE_ARENA_H_
#include <stddef.h>

void *oa_malloc(size_t n);
void *oa_calloc(size_t k, size_t n);
void *oa_realloc(void *p, size_t n);
void oa_free(void *p);
/* arena control (per calling thread) */
void oa_begin(size_t bytes);     /* allocate the thread's arena and switch it on */
void oa_reset(void);             /* forget everything allocated from it */
void oa_end(void);               /* switch off and release */
int oa_suspend(void);            /* switch off (allocations that must outlive the solve); returns the old state */
void oa_resume(int state);

#ifndef PQP_ORACLE_ARENA_IMPL
#define malloc(n) oa_malloc(n)
#define calloc(k, n) oa_calloc(k, n)
#define realloc(p, n) oa_realloc(p, n)
#define free(p) oa_free(p)
#endif
#endif
