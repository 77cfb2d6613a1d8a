/*
 * bcnn_cfg.c -- bcnn_load_net: build a net from a .cfg / .conf file, then load its weights.
 *
 * Accepts what jnbraun/bcnn accepts (src/bcnn_net.c:683-1218 with the ini reader of
 * src/bh/inc/bh/bh_ini.h:59-161), for the layer kinds this library has:
 *   * ini rules: blanks, tabs and newlines are removed from every line before anything else;
 *     a line starting with '[' opens a section (the whole stripped line is its name); lines
 *     starting with '#', ';', '!' and empty lines are skipped; every other line must split into
 *     exactly two tokens at '='. The first section must be [net] or [network] and hold keys.
 *   * [net] keys go to bcnn_net_set_param (shape, batch, solver); unknown keys are ignored.
 *   * two dialects, chosen like the reference by the extension of `model_path`: bcnn (explicit
 *     src= / dst= names, pad= is the padding in pixels) and Darknet when the model is a
 *     *.weights file (sections are chained implicitly as lid<i-1> -> lid<i>, pad=1 means
 *     size/2, route `layers=` and shortcut `from=` are section-relative).
 * Organisation: the file is read into a section list; per section one `layer_desc` is filled
 * from key tables (integer / float fields by offset, enumerations by word lists) and handed to
 * the matching bcnn_add_*_layer constructor.
 *
 * Not built here (the constructors do not exist on this path): [deconv], [lrn], [dropout]; a cfg
 * using them fails with BCNN_INVALID_PARAMETER and a log line naming the section. One stricter
 * check than the reference: a key line before any section is an error (the reference
 * dereferences NULL there).
 */
#include <ctype.h>
#include <stddef.h>

#include "bcnn_net.h"
#include <bcnn_b200_net.h>

/* ---------------------------------------------------------------- ini file -> sections */
typedef struct { char *name, *val; } cfg_key;
typedef struct { char *name; cfg_key *keys; int num_keys; } cfg_section;
typedef struct { cfg_section *sections; int num_sections; } cfg_file;

static void cfg_free(cfg_file *cfg) {
    for (int i = 0; i < cfg->num_sections; ++i) {
        for (int j = 0; j < cfg->sections[i].num_keys; ++j) {
            free(cfg->sections[i].keys[j].name);
            free(cfg->sections[i].keys[j].val);
        }
        free(cfg->sections[i].keys);
        free(cfg->sections[i].name);
    }
    free(cfg->sections);
    memset(cfg, 0, sizeof(*cfg));
}

static void strip_blanks(char *s) { /* bh_strstrip: ' ', '\t', '\n' anywhere in the line */
    char *out = s;
    for (; *s; ++s)
        if (*s != ' ' && *s != '\t' && *s != '\n') *out++ = *s;
    *out = '\0';
}

static char *read_text(const char *path) {
    FILE *fp = fopen(path, "rb");
    if (!fp) return NULL;
    fseek(fp, 0, SEEK_END);
    long size = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    char *text = size >= 0 ? (char *)malloc((size_t)size + 1) : NULL;
    if (text) {
        size_t got = fread(text, 1, (size_t)size, fp);
        text[got] = '\0';
    }
    fclose(fp);
    return text;
}

/* Returns 0, or -1 with a message in `why`. */
static int cfg_read(const char *path, cfg_file *cfg, char *why, size_t why_len) {
    memset(cfg, 0, sizeof(*cfg));
    char *text = read_text(path);
    if (!text) {
        snprintf(why, why_len, "Could not open file: %s", path);
        return -1;
    }
    int rc = 0;
    for (char *line = text, *next; rc == 0 && line; line = next) {
        next = strchr(line, '\n');
        if (next) *next++ = '\0';
        strip_blanks(line);
        const char first = line[0];
        if (first == '\0' || first == '#' || first == ';' || first == '!') continue;
        if (first == '[') {
            cfg_section *grown = (cfg_section *)realloc(
                cfg->sections, (size_t)(cfg->num_sections + 1) * sizeof(cfg_section));
            if (!grown) { rc = -1; snprintf(why, why_len, "Failed allocation"); break; }
            cfg->sections = grown;
            memset(&grown[cfg->num_sections], 0, sizeof(cfg_section));
            grown[cfg->num_sections++].name = bcnn_strdup_(line);
            continue;
        }
        char *eq = strchr(line, '=');
        if (!eq || strchr(eq + 1, '=')) { /* bh_strsplit(line, '=') must give exactly 2 tokens */
            snprintf(why, why_len, "Invalid key section %s", line);
            rc = -1;
            break;
        }
        if (cfg->num_sections == 0) {
            snprintf(why, why_len, "No valid section for key %s", line);
            rc = -1;
            break;
        }
        cfg_section *sec = &cfg->sections[cfg->num_sections - 1];
        cfg_key *keys = (cfg_key *)realloc(sec->keys, (size_t)(sec->num_keys + 1) * sizeof(cfg_key));
        if (!keys) { rc = -1; snprintf(why, why_len, "Failed allocation"); break; }
        sec->keys = keys;
        *eq = '\0';
        keys[sec->num_keys].name = bcnn_strdup_(line);
        keys[sec->num_keys].val = bcnn_strdup_(eq + 1);
        ++sec->num_keys;
    }
    free(text);
    if (rc != 0) cfg_free(cfg);
    return rc;
}

/* ---------------------------------------------------------------- one layer's parameters */
enum { CFG_MAX_SRCS = 16, CFG_MAX_ANCHORS = 64 };

typedef struct {
    int stride, pad, n_filts, size, outputs, num_groups, batchnorm;
    int in_w, in_h, in_c;
    int num_anchors, boxes_per_cell, num_classes, num_coords;
    float alpha, beta, k, rate;
    int padding_type, a, init, cost, loss; /* enums, stored as int for the tables */
    char *src_id[CFG_MAX_SRCS];
    int num_srcs;
    char *dst_id;
    int anchors_mask[CFG_MAX_ANCHORS];
    float anchors[2 * CFG_MAX_ANCHORS];
    int num_anchor_values;
} layer_desc;

static void layer_desc_clear(layer_desc *d) {
    for (int i = 0; i < d->num_srcs; ++i) free(d->src_id[i]);
    free(d->dst_id);
    memset(d, 0, sizeof(*d));
    /* defaults of bcnn_layer_param_reset (reference :716-755) */
    d->stride = 1;
    d->n_filts = 1;
    d->size = 3;
    d->num_groups = 1;
    d->num_coords = 4;
    d->rate = 1.0f;
    d->padding_type = BCNN_PADDING_SAME;
    d->a = BCNN_ACT_NONE;
    d->init = BCNN_FILLER_XAVIER;
    d->cost = BCNN_METRIC_SSE;
    d->loss = BCNN_LOSS_EUCLIDEAN;
}

static void set_srcs(layer_desc *d, int count, const char *const *names) {
    for (int i = 0; i < d->num_srcs; ++i) free(d->src_id[i]);
    d->num_srcs = count > CFG_MAX_SRCS ? CFG_MAX_SRCS : count;
    for (int i = 0; i < d->num_srcs; ++i) d->src_id[i] = bcnn_strdup_(names[i]);
}

/* Comma-separated list, tokens handed to `each` (at most `max`); returns the token count. */
static int for_each_token(const char *val, int max, void (*each)(const char *, int, void *),
                          void *arg) {
    char *copy = bcnn_strdup_(val);
    int n = 0;
    for (char *tok = copy, *next; tok; tok = next) {
        next = strchr(tok, ',');
        if (next) *next++ = '\0';
        if (n < max) each(tok, n, arg);
        ++n;
    }
    free(copy);
    return n > max ? max : n;
}

typedef struct { layer_desc *d; int section; char names[CFG_MAX_SRCS][256]; } route_ctx;

static void take_anchor(const char *tok, int i, void *arg) { ((layer_desc *)arg)->anchors[i] = (float)atof(tok); }
static void take_mask(const char *tok, int i, void *arg) { ((layer_desc *)arg)->anchors_mask[i] = atoi(tok); }
static void take_src(const char *tok, int i, void *arg) {
    route_ctx *r = (route_ctx *)arg;
    snprintf(r->names[i], sizeof(r->names[i]), "%s", tok);
    (void)i;
}
/* Darknet layer reference -> tensor name: absolute index l names section l+1, a negative one is
 * relative to the current section (reference :936-943). */
static void darknet_layer_name(char *out, size_t len, int section, int l) {
    snprintf(out, len, "lid%d", l >= 0 ? l + 1 : section + l);
}
static void take_route(const char *tok, int i, void *arg) {
    route_ctx *r = (route_ctx *)arg;
    darknet_layer_name(r->names[i], sizeof(r->names[i]), r->section, atoi(tok));
}

typedef struct { const char *word; int value; } word_value;

static int lookup_word(const word_value *table, int count, const char *word, int fallback,
                       bcnn_net *net, const char *what) {
    for (int i = 0; i < count; ++i)
        if (!strcmp(word, table[i].word)) return table[i].value;
    if (what) BCNN_WARNING(net->log_ctx, "Unknown %s %s\n", what, word);
    return fallback;
}

#define COUNT(a) ((int)(sizeof(a) / sizeof((a)[0])))

static void layer_desc_set(bcnn_net *net, int section, layer_desc *d, const char *name,
                           const char *val, int darknet) {
    /* plain integer fields (atoi), by name; aliases share an offset */
    static const struct { const char *key; size_t offset; } int_fields[] = {
        {"filters", offsetof(layer_desc, n_filts)},       {"size", offsetof(layer_desc, size)},
        {"stride", offsetof(layer_desc, stride)},         {"num_groups", offsetof(layer_desc, num_groups)},
        {"groups", offsetof(layer_desc, num_groups)},     {"boxes_per_cell", offsetof(layer_desc, boxes_per_cell)},
        {"num_anchors", offsetof(layer_desc, num_anchors)}, {"num", offsetof(layer_desc, num_anchors)},
        {"num_classes", offsetof(layer_desc, num_classes)}, {"classes", offsetof(layer_desc, num_classes)},
        {"num_coords", offsetof(layer_desc, num_coords)}, {"w", offsetof(layer_desc, in_w)},
        {"h", offsetof(layer_desc, in_h)},                {"c", offsetof(layer_desc, in_c)},
        {"bn", offsetof(layer_desc, batchnorm)},          {"batchnorm", offsetof(layer_desc, batchnorm)},
        {"batch_normalize", offsetof(layer_desc, batchnorm)}, {"output", offsetof(layer_desc, outputs)}};
    static const word_value activations[] = {
        {"relu", BCNN_ACT_RELU},       {"tanh", BCNN_ACT_TANH},   {"ramp", BCNN_ACT_RAMP},
        {"clamp", BCNN_ACT_CLAMP},     {"softplus", BCNN_ACT_SOFTPLUS}, {"leaky_relu", BCNN_ACT_LRELU},
        {"lrelu", BCNN_ACT_LRELU},     {"leaky", BCNN_ACT_LRELU}, {"prelu", BCNN_ACT_PRELU},
        {"abs", BCNN_ACT_ABS},         {"none", BCNN_ACT_NONE},   {"linear", BCNN_ACT_NONE}};
    static const word_value fillers[] = {{"xavier", BCNN_FILLER_XAVIER}, {"msra", BCNN_FILLER_MSRA}};
    static const word_value metrics[] = {
        {"error", BCNN_METRIC_ERROR_RATE}, {"logloss", BCNN_METRIC_LOGLOSS}, {"sse", BCNN_METRIC_SSE},
        {"mse", BCNN_METRIC_MSE},          {"crps", BCNN_METRIC_CRPS},       {"dice", BCNN_METRIC_DICE}};
    static const word_value losses[] = {{"l2", BCNN_LOSS_EUCLIDEAN},
                                        {"euclidean", BCNN_LOSS_EUCLIDEAN},
                                        {"lifted_struct_similarity", BCNN_LOSS_LIFTED_STRUCT}};
    static const word_value paddings[] = {{"same", BCNN_PADDING_SAME},
                                          {"valid", BCNN_PADDING_VALID},
                                          {"caffe", BCNN_PADDING_CAFFE}};
    for (int i = 0; i < COUNT(int_fields); ++i)
        if (!strcmp(name, int_fields[i].key)) {
            *(int *)((char *)d + int_fields[i].offset) = atoi(val);
            return;
        }
    if (!strcmp(name, "dropout_rate") || !strcmp(name, "rate")) {
        d->rate = (float)atof(val);
    } else if (!strcmp(name, "alpha")) { /* the reference reads these three with atoi */
        d->alpha = (float)atoi(val);
    } else if (!strcmp(name, "beta")) {
        d->beta = (float)atoi(val);
    } else if (!strcmp(name, "k")) {
        d->k = (float)atoi(val);
    } else if (!strcmp(name, "padding")) { /* Darknet max-pool padding; ignored in bcnn files */
        if (darknet) {
            d->pad = atoi(val);
            d->padding_type = d->pad ? BCNN_PADDING_SAME : BCNN_PADDING_VALID;
        }
    } else if (!strcmp(name, "pad")) { /* bcnn: pixels; Darknet: flag for size/2 (size so far) */
        d->pad = darknet ? (atoi(val) ? d->size / 2 : 0) : atoi(val);
    } else if (!strcmp(name, "anchors")) {
        d->num_anchor_values = for_each_token(val, 2 * CFG_MAX_ANCHORS, take_anchor, d);
    } else if (!strcmp(name, "anchors_mask") || !strcmp(name, "mask")) {
        d->boxes_per_cell = for_each_token(val, CFG_MAX_ANCHORS, take_mask, d);
    } else if (!strcmp(name, "src")) {
        route_ctx r = {d, section, {{0}}};
        int n = for_each_token(val, CFG_MAX_SRCS, take_src, &r);
        const char *names[CFG_MAX_SRCS];
        for (int i = 0; i < n; ++i) names[i] = r.names[i];
        set_srcs(d, n, names);
    } else if (!strcmp(name, "dst")) {
        free(d->dst_id);
        d->dst_id = bcnn_strdup_(val);
    } else if (!strcmp(name, "padding_type")) {
        d->padding_type = lookup_word(paddings, COUNT(paddings), val, d->padding_type, net, NULL);
    } else if (!strcmp(name, "function") || !strcmp(name, "activation")) {
        d->a = lookup_word(activations, COUNT(activations), val, BCNN_ACT_RELU, net,
                           "activation type (going with ReLU)");
    } else if (!strcmp(name, "init")) {
        d->init = lookup_word(fillers, COUNT(fillers), val, BCNN_FILLER_XAVIER, net,
                              "init type (going with xavier)");
    } else if (!strcmp(name, "metric")) {
        d->cost = lookup_word(metrics, COUNT(metrics), val, BCNN_METRIC_SSE, net,
                              "cost metric (going with sse)");
    } else if (!strcmp(name, "loss")) {
        d->loss = lookup_word(losses, COUNT(losses), val, BCNN_LOSS_EUCLIDEAN, net,
                              "loss (going with euclidean)");
    } else if (!strcmp(name, "layers")) { /* Darknet [route] */
        route_ctx r = {d, section, {{0}}};
        int n = for_each_token(val, CFG_MAX_SRCS, take_route, &r);
        const char *names[CFG_MAX_SRCS];
        for (int i = 0; i < n; ++i) names[i] = r.names[i];
        set_srcs(d, n, names);
    } else if (!strcmp(name, "from")) { /* Darknet [shortcut]: previous section + the named one */
        char a[32], b[32];
        snprintf(a, sizeof(a), "lid%d", section - 1);
        darknet_layer_name(b, sizeof(b), section, atoi(val));
        const char *names[2] = {a, b};
        set_srcs(d, 2, names);
    }
    /* anything else: ignored, like the reference */
}

/* ---------------------------------------------------------------- section -> constructor */
typedef enum {
    SEC_INPUT, SEC_CONV, SEC_DEPTHWISE, SEC_ACTIVATION, SEC_BATCHNORM, SEC_FULLC, SEC_SOFTMAX,
    SEC_MAXPOOL, SEC_AVGPOOL, SEC_UPSAMPLE, SEC_CONCAT, SEC_ELTWISE, SEC_YOLO, SEC_COST,
    SEC_UNSUPPORTED, SEC_UNKNOWN
} section_kind;

static section_kind kind_of(const char *name) {
    static const struct { const char *name; section_kind kind; } names[] = {
        {"[input]", SEC_INPUT}, {"[conv]", SEC_CONV}, {"[convolutional]", SEC_CONV},
        {"[depthwise-conv]", SEC_DEPTHWISE}, {"[dw-conv]", SEC_DEPTHWISE},
        {"[activation]", SEC_ACTIVATION}, {"[nl]", SEC_ACTIVATION},
        {"[batchnorm]", SEC_BATCHNORM}, {"[bn]", SEC_BATCHNORM},
        {"[connected]", SEC_FULLC}, {"[fullconnected]", SEC_FULLC}, {"[fc]", SEC_FULLC},
        {"[ip]", SEC_FULLC}, {"[softmax]", SEC_SOFTMAX}, {"[max]", SEC_MAXPOOL},
        {"[maxpool]", SEC_MAXPOOL}, {"[avgpool]", SEC_AVGPOOL}, {"[upsample]", SEC_UPSAMPLE},
        {"[concat]", SEC_CONCAT}, {"[route]", SEC_CONCAT}, {"[eltwise]", SEC_ELTWISE},
        {"[shortcut]", SEC_ELTWISE}, {"[yolo]", SEC_YOLO}, {"[cost]", SEC_COST},
        {"[deconv]", SEC_UNSUPPORTED}, {"[deconvolutional]", SEC_UNSUPPORTED},
        {"[lrn]", SEC_UNSUPPORTED}, {"[dropout]", SEC_UNSUPPORTED}};
    for (int i = 0; i < COUNT(names); ++i)
        if (!strcmp(name, names[i].name)) return names[i].kind;
    return SEC_UNKNOWN;
}

static bcnn_status add_layer(bcnn_net *net, const char *name, const layer_desc *d) {
    const section_kind kind = kind_of(name);
    BCNN_CHECK_AND_LOG(net->log_ctx, kind != SEC_UNKNOWN, BCNN_INVALID_PARAMETER,
                       "Unknown Layer %s\n", name);
    BCNN_CHECK_AND_LOG(net->log_ctx, kind != SEC_UNSUPPORTED, BCNN_INVALID_PARAMETER,
                       "Layer %s is not available on the B200 path\n", name);
    BCNN_CHECK_AND_LOG(net->log_ctx, d->num_srcs > 0 && d->src_id[0], BCNN_INVALID_PARAMETER,
                       "Invalid input node name. Hint: Are you sure that 'src' field is "
                       "correctly setup?\n");
    const int needs_dst = kind != SEC_INPUT && kind != SEC_ACTIVATION;
    BCNN_CHECK_AND_LOG(net->log_ctx, !needs_dst || d->dst_id, BCNN_INVALID_PARAMETER,
                       "Invalid output node name. Hint: Are you sure that 'dst' field is "
                       "correctly setup?\n");
    const char *src = d->src_id[0], *dst = d->dst_id;
    switch (kind) {
        case SEC_INPUT:
            return bcnn_add_input(net, d->in_w, d->in_h, d->in_c, src);
        case SEC_CONV:
            return bcnn_add_convolutional_layer(net, d->n_filts, d->size, d->stride, d->pad,
                                                d->num_groups, d->batchnorm,
                                                (bcnn_filler_type)d->init, (bcnn_activation)d->a,
                                                0, src, dst);
        case SEC_DEPTHWISE:
            return bcnn_add_depthwise_conv_layer(net, d->size, d->stride, d->pad, 0,
                                                 (bcnn_filler_type)d->init, (bcnn_activation)d->a,
                                                 src, dst);
        case SEC_ACTIVATION:
            return bcnn_add_activation_layer(net, (bcnn_activation)d->a, src);
        case SEC_BATCHNORM:
            return bcnn_add_batchnorm_layer(net, src, dst);
        case SEC_FULLC:
            return bcnn_add_fullc_layer(net, d->outputs, (bcnn_filler_type)d->init,
                                        (bcnn_activation)d->a, 0, src, dst);
        case SEC_SOFTMAX:
            return bcnn_add_softmax_layer(net, src, dst);
        case SEC_MAXPOOL:
            return bcnn_add_maxpool_layer(net, d->size, d->stride, (bcnn_padding)d->padding_type,
                                          src, dst);
        case SEC_AVGPOOL:
            return bcnn_add_avgpool_layer(net, src, dst);
        case SEC_UPSAMPLE:
            return bcnn_add_upsample_layer(net, d->stride, src, dst);
        case SEC_CONCAT:
            return bcnn_add_concat_layer(net, d->num_srcs, (char *const *)d->src_id, dst);
        case SEC_ELTWISE:
            BCNN_CHECK_AND_LOG(net->log_ctx, d->num_srcs >= 2, BCNN_INVALID_PARAMETER,
                               "Eltwise layer needs two inputs\n");
            return bcnn_add_eltwise_layer(net, (bcnn_activation)d->a, d->src_id[0], d->src_id[1],
                                          dst);
        case SEC_YOLO:
            BCNN_CHECK_AND_LOG(net->log_ctx, d->num_anchors <= CFG_MAX_ANCHORS, BCNN_INVALID_PARAMETER,
                               "Yolo layer: more than %d anchors\n", CFG_MAX_ANCHORS);
            return bcnn_add_yolo_layer(net, d->boxes_per_cell, d->num_classes, d->num_coords,
                                       d->num_anchors, (int *)d->anchors_mask,
                                       d->num_anchor_values ? (float *)d->anchors : NULL, src, dst);
        case SEC_COST:
            return bcnn_add_cost_layer(net, (bcnn_loss)d->loss, (bcnn_loss_metric)d->cost, 1.0f,
                                       src, "label", dst);
        default:
            return BCNN_INVALID_PARAMETER;
    }
}

static int has_suffix(const char *path, const char *ext) {
    const char *dot = strrchr(path, '.');
    return dot && !strcmp(dot + 1, ext);
}

bcnn_status bcnn_load_net(bcnn_net *net, const char *config_path, const char *model_path) {
    BCNN_CHECK_AND_LOG(net->log_ctx, config_path != NULL, BCNN_INVALID_PARAMETER,
                       "bcnn_load_net: no config file\n");
    int darknet = 0, onnx = 0;
    if (model_path) {
        BCNN_CHECK_AND_LOG(net->log_ctx, strchr(model_path, '.') != NULL, BCNN_INVALID_DATA,
                           "File %s needs to have an extension (.bcnnmodel OR .onnx OR .weights)\n",
                           model_path);
        darknet = has_suffix(model_path, "weights");
        onnx = has_suffix(model_path, "onnx");
    }
    if (!onnx) { /* the reference skips the cfg for an .onnx model and then fails in the loader */
        cfg_file cfg;
        char why[512];
        if (cfg_read(config_path, &cfg, why, sizeof(why)) != 0) {
            bcnn_log(net->log_ctx, BCNN_LOG_ERROR, "%s\n", why);
            return BCNN_INVALID_PARAMETER;
        }
        bcnn_status st = BCNN_SUCCESS;
        if (cfg.num_sections == 0) {
            bcnn_log(net->log_ctx, BCNN_LOG_ERROR, "Empty config file %s\n", config_path);
            st = BCNN_INVALID_PARAMETER;
        } else if (strcmp(cfg.sections[0].name, "[net]") && strcmp(cfg.sections[0].name, "[network]")) {
            bcnn_log(net->log_ctx, BCNN_LOG_ERROR,
                     "Invalid config file %s: First section must be [net] or [network]\n",
                     config_path);
            st = BCNN_INVALID_PARAMETER;
        } else if (cfg.sections[0].num_keys == 0) {
            bcnn_log(net->log_ctx, BCNN_LOG_ERROR, "Invalid config file %s: empty section [net]\n",
                     config_path);
            st = BCNN_INVALID_PARAMETER;
        }
        if (st == BCNN_SUCCESS) {
            for (int j = 0; j < cfg.sections[0].num_keys; ++j)
                bcnn_net_set_param(net, cfg.sections[0].keys[j].name, cfg.sections[0].keys[j].val);
            layer_desc d;
            memset(&d, 0, sizeof(d));
            for (int i = 1; i < cfg.num_sections && st == BCNN_SUCCESS; ++i) {
                layer_desc_clear(&d);
                const cfg_section *sec = &cfg.sections[i];
                for (int j = 0; j < sec->num_keys; ++j)
                    layer_desc_set(net, i, &d, sec->keys[j].name, sec->keys[j].val, darknet);
                if (darknet) { /* implicit chain: section i reads lid<i-1>, writes lid<i> */
                    char lid[32];
                    if (d.num_srcs == 0) {
                        snprintf(lid, sizeof(lid), "lid%d", i - 1);
                        const char *names[1] = {lid};
                        set_srcs(&d, 1, names);
                    }
                    if (!d.dst_id) {
                        snprintf(lid, sizeof(lid), "lid%d", i);
                        d.dst_id = bcnn_strdup_(lid);
                    }
                }
                if (add_layer(net, sec->name, &d) != BCNN_SUCCESS) st = BCNN_INVALID_PARAMETER;
            }
            layer_desc_clear(&d);
        }
        cfg_free(&cfg);
        BCNN_CHECK_STATUS(st);
    }
    if (model_path) {
        BCNN_INFO(net->log_ctx, "Loading pre-trained model %s\n", model_path);
        BCNN_CHECK_STATUS(bcnn_load_weights(net, model_path));
    }
    return BCNN_SUCCESS;
}
