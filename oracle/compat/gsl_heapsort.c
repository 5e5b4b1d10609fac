/*
 * oracle/compat/gsl_heapsort.c -- gsl_heapsort_index, the one GSL routine on the hot path
 * (sort.c:192).  TEST INFRASTRUCTURE ONLY.  Restated from the published algorithm (index
 * heapsort with sift-down: build the heap from k = N/2 down to 0, then swap the root with the
 * last element and sift), so ties come out in the same (unstable) order.
 */
#include <stddef.h>
#include "gsl/gsl_heapsort.h"

typedef int (*cmp_fn)(const void *, const void *);

static void sift(size_t *p, const char *base, size_t size, size_t last,
                 size_t k, cmp_fn cmp)
{
    const size_t moving = p[k];

    while (k <= last / 2) {
        size_t child = 2 * k;

        if (child < last &&
            cmp(base + p[child] * size, base + p[child + 1] * size) < 0)
            child++;

        if (cmp(base + moving * size, base + p[child] * size) >= 0)
            break;

        p[k] = p[child];
        k = child;
    }
    p[k] = moving;
}

int gsl_heapsort_index(size_t *p, const void *array, size_t count, size_t size,
                       cmp_fn cmp)
{
    if (count == 0)
        return 0;

    for (size_t i = 0; i < count; i++)
        p[i] = i;

    size_t last = count - 1;

    for (size_t k = last / 2 + 1; k-- > 0;)
        sift(p, array, size, last, k, cmp);

    while (last > 0) {
        size_t t = p[0];
        p[0] = p[last];
        p[last] = t;
        last--;
        sift(p, array, size, last, 0, cmp);
    }
    return 0;
}

