/* utils.c -- timers used by the progress lines (reference utils.c:1-18). */
#include <sys/resource.h>
#include <sys/time.h>
#include "bfc.h"

double cputime(void)
{
	struct rusage r;
	getrusage(RUSAGE_SELF, &r);
	return (double)(r.ru_utime.tv_sec + r.ru_stime.tv_sec) + 1e-6 * (double)(r.ru_utime.tv_usec + r.ru_stime.tv_usec);
}

double realtime(void)
{
	struct timeval tv;
	gettimeofday(&tv, 0);
	return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
}
