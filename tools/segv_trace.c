/* LD_PRELOAD helper: print a backtrace on SIGSEGV (debugging aid for the GPU box, no gdb there) */
#define _GNU_SOURCE
#include <execinfo.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
static void handler(int sig)
{
    void * bt[64];
    int n = backtrace(bt, 64);
    fprintf(stderr, "signal %d, backtrace:\n", sig);
    backtrace_symbols_fd(bt, n, 2);
    _exit(139);
}
__attribute__((constructor)) static void init(void) { signal(SIGSEGV, handler); signal(SIGABRT, handler); }
