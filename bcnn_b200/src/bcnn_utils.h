/*
 * bcnn_utils.h -- logging, status-check macros and the CUDA fail-fast convention
 * of the host runtime. Same macro names and behaviour as jnbraun/bcnn's
 * src/bcnn_utils.h:70-107,174-180 (log + early return at construction time;
 * device failures print to stderr and exit), so layer code reads the same.
 */
#ifndef BCNN_UTILS_H
#define BCNN_UTILS_H

#include <bcnn/bcnn.h>
#include <bcnn_b200.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    bcnn_log_callback fct;
    bcnn_log_level lvl;
} bcnn_log_context;

void bcnn_log(bcnn_log_context ctx, bcnn_log_level level, const char *fmt, ...);

#define BCNN_CHECK(exp, err) \
    do { if (!(exp)) return (err); } while (0)
#define BCNN_CHECK_AND_LOG(ctx, exp, err, fmt, ...)                          \
    do { if (!(exp)) { bcnn_log((ctx), BCNN_LOG_ERROR, (fmt), ##__VA_ARGS__); \
                       return (err); } } while (0)
#define BCNN_CHECK_STATUS(s) \
    do { bcnn_status ret_ = (s); if (ret_ != BCNN_SUCCESS) return ret_; } while (0)
#define BCNN_ERROR(ctx, err, fmt, ...) \
    do { bcnn_log((ctx), BCNN_LOG_ERROR, (fmt), ##__VA_ARGS__); return (err); } while (0)
#define BCNN_INFO(ctx, fmt, ...) bcnn_log((ctx), BCNN_LOG_INFO, (fmt), ##__VA_ARGS__)
#define BCNN_WARNING(ctx, fmt, ...) bcnn_log((ctx), BCNN_LOG_WARNING, (fmt), ##__VA_ARGS__)

/* Kernel launchers return a cudaError_t as int: print and exit, like the reference. */
#define bcnn_cuda_check(RET)                                                          \
    do { int r_ = (RET);                                                              \
         if (r_ != 0) {                                                               \
             fprintf(stderr, "[ERROR] [CUDA] %s (%s:%d)\n", bcnn_b200_error_string(r_), \
                     __FILE__, __LINE__);                                             \
             exit(r_); } } while (0)

const char *bcnn_act2str(bcnn_activation a);
char *bcnn_strdup_(const char *s);

static inline int bcnn_min_i(int a, int b) { return a < b ? a : b; }
static inline int bcnn_max_i(int a, int b) { return a > b ? a : b; }

#ifdef __cplusplus
}
#endif
#endif /* BCNN_UTILS_H */
