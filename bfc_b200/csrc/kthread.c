/* kthread.c -- the two host-side schedulers the phase drivers use
 * (interface: reference bfc.h:42-43; semantics: reference kthread.c).
 *
 * kt_for      runs func(data, i, tid) for i in [0, n) on n_threads threads; tid < n_threads
 *             identifies the worker (per-thread scratch).  With n_threads == 1 the items
 *             run in order 0, 1, 2, ... on the calling thread's behalf.
 * kt_pipeline runs batches through n_steps steps.  Step s of batch b starts only after
 *             step s of batch b-1 has finished, so a step never overlaps itself and
 *             batches leave every step in order (reference kthread.c:92-102); with two
 *             workers reading batch b+1 overlaps computing batch b.  Step 0 returning
 *             NULL ends the input.
 */
#include <pthread.h>
#include <stdlib.h>
#include <stdint.h>
#include "bfc.h"

/* ---------------------------------------------------------------- kt_for */

typedef struct {
	void (*func)(void*, long, int);
	void *data;
	long n;
	int n_threads;
	volatile long next; /* dynamic schedule: items are claimed one at a time */
} for_shared_t;

typedef struct { for_shared_t *sh; int tid; } for_worker_t;

static void *for_worker(void *arg)
{
	for_worker_t *w = (for_worker_t*)arg;
	for (;;) {
		long i = __sync_fetch_and_add(&w->sh->next, 1);
		if (i >= w->sh->n) break;
		w->sh->func(w->sh->data, i, w->tid);
	}
	return 0;
}

void kt_for(int n_threads, void (*func)(void*, long, int), void *data, long n)
{
	for_shared_t sh;
	int i;
	if (n_threads < 1) n_threads = 1;
	sh.func = func, sh.data = data, sh.n = n, sh.n_threads = n_threads, sh.next = 0;
	if (n_threads == 1) {
		long j;
		for (j = 0; j < n; ++j) func(data, j, 0);
		return;
	}
	{
		pthread_t *tid = (pthread_t*)malloc(n_threads * sizeof(pthread_t));
		for_worker_t *w = (for_worker_t*)malloc(n_threads * sizeof(for_worker_t));
		for (i = 0; i < n_threads; ++i) w[i].sh = &sh, w[i].tid = i;
		for (i = 0; i < n_threads; ++i) pthread_create(&tid[i], 0, for_worker, &w[i]);
		for (i = 0; i < n_threads; ++i) pthread_join(tid[i], 0);
		free(w); free(tid);
	}
}

/* ---------------------------------------------------------------- kt_pipeline */

typedef struct {
	void *(*func)(void*, int, void*);
	void *shared;
	int n_steps;
	int64_t next_batch;      /* next batch index to hand out */
	int64_t *done;           /* done[s] = number of batches that have left step s */
	int eof;                 /* step 0 returned NULL */
	pthread_mutex_t mu;
	pthread_cond_t cv;
} pipe_t;

static void *pipe_worker(void *arg)
{
	pipe_t *p = (pipe_t*)arg;
	for (;;) {
		int64_t b;
		int s, alive = 1;
		void *data = 0;
		pthread_mutex_lock(&p->mu);
		b = p->next_batch++;
		pthread_mutex_unlock(&p->mu);
		for (s = 0; s < p->n_steps; ++s) {
			pthread_mutex_lock(&p->mu);
			while (p->done[s] != b) pthread_cond_wait(&p->cv, &p->mu); /* batch b-1 must have left step s */
			if (s == 0 && p->eof) alive = 0;
			pthread_mutex_unlock(&p->mu);
			if (alive) {
				data = p->func(p->shared, s, s ? data : 0);
				if (data == 0 && (s == 0 || s < p->n_steps - 1)) alive = 0; /* end of input / nothing to pass on */
			}
			pthread_mutex_lock(&p->mu);
			if (s == 0 && !alive) p->eof = 1;
			++p->done[s];
			pthread_cond_broadcast(&p->cv);
			pthread_mutex_unlock(&p->mu);
		}
		if (p->eof) break;
	}
	return 0;
}

void kt_pipeline(int n_threads, void *(*func)(void*, int, void*), void *shared_data, int n_steps)
{
	pipe_t p;
	pthread_t *tid;
	int i;
	if (n_threads < 1) n_threads = 1;
	p.func = func, p.shared = shared_data, p.n_steps = n_steps, p.next_batch = 0, p.eof = 0;
	p.done = (int64_t*)calloc(n_steps, sizeof(int64_t));
	pthread_mutex_init(&p.mu, 0);
	pthread_cond_init(&p.cv, 0);
	tid = (pthread_t*)malloc(n_threads * sizeof(pthread_t));
	for (i = 0; i < n_threads; ++i) pthread_create(&tid[i], 0, pipe_worker, &p);
	for (i = 0; i < n_threads; ++i) pthread_join(tid[i], 0);
	free(tid); free(p.done);
	pthread_mutex_destroy(&p.mu);
	pthread_cond_destroy(&p.cv);
}
