/*
 * oracle_bench.c -- frame-parallel timing harness around the C restatement (TEST INFRASTRUCTURE ONLY).
 * Used by bench.py's cpu_baseline leg with kind "port" when oracle/_ref is unavailable.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

size_t oracle_decode(uint16_t*, size_t, int, int, const uint8_t*, size_t);
size_t oracle_decode_legacy(uint16_t*, size_t, int, int, const uint8_t*, size_t);

typedef struct {
    int t, threads, type, nframes, width, height, iters;
    const uint8_t* const* ins;
    const size_t* lens;
    int64_t done;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    size_t cap = (size_t)j->width * j->height;
    uint16_t* out = (uint16_t*)malloc(cap * 2 + 128);
    memset(out, 1, cap * 2 + 128);
    for (int i = 0; i < j->iters; i++)
        for (int f = j->t; f < j->nframes; f += j->threads) {
            size_t r = j->type == 6 ? oracle_decode_legacy(out, cap, j->width, j->height, j->ins[f], j->lens[f])
                                    : oracle_decode(out, cap, j->width, j->height, j->ins[f], j->lens[f]);
            if (r) j->done++;
        }
    free(out);
    return NULL;
}

double oracle_bench_mt(int compression_type, const uint8_t* const* ins, const size_t* lens, int nframes,
                       int width, int height, int threads, int iters, int64_t* frames_done) {
    if (threads < 1) threads = 1;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = 0; t < threads; t++) {
        jobs[t] = (job_t){t, threads, compression_type, nframes, width, height, iters, ins, lens, 0};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    int64_t done = 0;
    for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); done += jobs[t].done; }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (frames_done) *frames_done = done;
    free(th);
    free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
