/* TEST INFRASTRUCTURE ONLY: LD_PRELOAD shim that counts calls to malloc (operator new lands here too). */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stddef.h>
static void* (*real_malloc)(size_t) = 0;
static long g_mallocs = 0;
void* malloc(size_t n) {
  if (!real_malloc) real_malloc = (void* (*)(size_t))dlsym(RTLD_NEXT, "malloc");
  __atomic_fetch_add(&g_mallocs, 1, __ATOMIC_RELAXED);
  return real_malloc(n);
}
long malloc_count_get(void) { return g_mallocs; }
