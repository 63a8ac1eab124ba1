// timer.h -- tic/toc stopwatch with the interface of the reference's include/timer.h:24-45
// (C++ linkage, like the reference: timer_create() overloads the POSIX function of that name)
#ifndef __B2_TIMER_H__
#define __B2_TIMER_H__

typedef struct timer_s * timer;

timer timer_create();
void  timer_destroy(timer _q);
void  timer_tic(timer _q);            // start the stopwatch
float timer_toc(timer _q);            // seconds since the last tic

#endif
