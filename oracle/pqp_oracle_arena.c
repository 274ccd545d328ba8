/* pqp_oracle_arena.c -- see pqp_oracle_arena.h (CPU ORACLE, test infrastructure). */
#define PQP_ORACLE_ARENA_IMPL
#include "pqp_oracle_arena.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { char *base; size_t cap, used; int active; } oa_arena;
static __thread oa_arena oa_tls = {0, 0, 0, 0};

static int oa_owns(const void *p) {
    return oa_tls.base && (const char *)p >= oa_tls.base && (const char *)p < oa_tls.base + oa_tls.cap;
}
void *oa_malloc(size_t n) {
    if (oa_tls.active) {
        const size_t need = (n + 63) & ~(size_t)63;
        if (oa_tls.used + need <= oa_tls.cap) {
            void *p = oa_tls.base + oa_tls.used;
            oa_tls.used += need ? need : 64;
            return p;
        }
    }
    return malloc(n);   /* arena off or exhausted */
}
void *oa_calloc(size_t k, size_t n) {
    if (oa_tls.active) {
        void *p = oa_malloc(k * n);
        if (p) memset(p, 0, k * n);
        return p;
    }
    return calloc(k, n);
}
void *oa_realloc(void *p, size_t n) {
    if (p && oa_owns(p)) {   /* sizes are not tracked: only ever grows small index vectors */
        void *q = oa_malloc(n);
        if (q) memcpy(q, p, n);   /* may over-read inside the arena block: harmless */
        return q;
    }
    if (!p) return oa_malloc(n);
    return realloc(p, n);
}
void oa_free(void *p) {
    if (!p || oa_owns(p)) return;
    free(p);
}
void oa_begin(size_t bytes) {
    if (oa_tls.base) free(oa_tls.base);
    oa_tls.base = (char *)aligned_alloc(64, (bytes + 63) & ~(size_t)63);
    oa_tls.cap = oa_tls.base ? ((bytes + 63) & ~(size_t)63) : 0;
    oa_tls.used = 0;
    oa_tls.active = oa_tls.base != NULL;
}
void oa_reset(void) { oa_tls.used = 0; }
void oa_end(void) {
    oa_tls.active = 0;
    free(oa_tls.base);
    oa_tls.base = NULL;
    oa_tls.cap = oa_tls.used = 0;
}
int oa_suspend(void) { const int s = oa_tls.active; oa_tls.active = 0; return s; }
void oa_resume(int state) { oa_tls.active = state; }
