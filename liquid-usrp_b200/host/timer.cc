// timer.cc -- stopwatch (reference: lib/timer.cc:40-84)
#include <stdlib.h>
#include <sys/time.h>
#include "timer.h"

struct timer_s {
    struct timeval t0;
    int running;
};

timer timer_create()
{
    timer q = (timer)calloc(1, sizeof(struct timer_s));
    return q;
}
void timer_destroy(timer q) { free(q); }
void timer_tic(timer q)
{
    gettimeofday(&q->t0, NULL);
    q->running = 1;
}
float timer_toc(timer q)
{
    if (!q->running) return 0.0f;
    struct timeval t1;
    gettimeofday(&t1, NULL);
    return (float)((double)(t1.tv_sec - q->t0.tv_sec) + 1e-6 * (double)(t1.tv_usec - q->t0.tv_usec));
}
