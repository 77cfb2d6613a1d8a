/* bcnn_utils.c -- logging helpers (behaviour of reference src/bcnn_utils.c:30-46). */
#include "bcnn_utils.h"

#include <stdarg.h>

void bcnn_log(bcnn_log_context ctx, bcnn_log_level level, const char *fmt, ...) {
    if (level < ctx.lvl) return;
    char msg[2048];
    va_list args;
    va_start(args, fmt);
    vsnprintf(msg, sizeof(msg), fmt, args);
    va_end(args);
    if (ctx.fct) {
        ctx.fct("%s", msg);
        return;
    }
    static const char *tag[] = {"[INFO]", "[WARNING]", "[ERROR]", ""};
    fprintf(stderr, "%s %s", tag[level > 3 ? 3 : level], msg);
}

const char *bcnn_act2str(bcnn_activation a) {
    static const char *names[] = {"None", "Tanh", "ReLU", "Ramp", "Softplus", "Leaky-ReLU",
                                  "AbsVal", "Clamp", "PReLU", "Logistic"};
    return ((int)a >= 0 && (int)a < 10) ? names[a] : "None";
}

char *bcnn_strdup_(const char *s) {
    size_t n = strlen(s) + 1;
    char *d = (char *)malloc(n);
    if (d) memcpy(d, s, n);
    return d;
}
