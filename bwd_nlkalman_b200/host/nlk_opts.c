/* nlk_opts.c -- see nlk_opts.h */
#include "nlk_opts.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void usage(const struct nlk_opt *opts, const char *usage_line, const char *descr)
{
    printf("Usage: %s\n", usage_line);
    if (descr) printf("%s\n", descr);
    printf("\n");
    printf("    -h, --help                show this help message and exit\n");
    for (const struct nlk_opt *o = opts; o->type != NLK_OPT_END; ++o) {
        if (o->type == NLK_OPT_GROUP) {
            printf("\n%s\n", o->long_name);
            continue;
        }
        char name[64];
        const char *meta = o->type == NLK_OPT_STRING ? "=<str>" : (o->type == NLK_OPT_INT ? "=<int>" : "=<flt>");
        if (o->short_name && o->long_name) snprintf(name, sizeof name, "-%c, --%s%s", o->short_name, o->long_name, meta);
        else if (o->long_name) snprintf(name, sizeof name, "--%s%s", o->long_name, meta);
        else snprintf(name, sizeof name, "-%c%s", o->short_name, meta);
        printf("    %-25s %s\n", name, o->help ? o->help : "");
    }
}

static void die(const char *what, const char *opt, int is_long, char sn)
{
    if (is_long) fprintf(stderr, "error: option `--%s` %s\n", opt, what);
    else fprintf(stderr, "error: option `-%c` %s\n", sn, what);
    exit(1);
}

static void set_value(const struct nlk_opt *o, const char *text, int is_long)
{
    char *end = NULL;
    switch (o->type) {
    case NLK_OPT_STRING:
        *(const char **)o->value = text;
        break;
    case NLK_OPT_INT:
        *(int *)o->value = (int)strtol(text, &end, 0);
        if (!*text || *end) die("expects an integer value", o->long_name, is_long, o->short_name);
        break;
    case NLK_OPT_FLOAT:
        *(float *)o->value = strtof(text, &end);
        if (!*text || *end) die("expects a numerical value", o->long_name, is_long, o->short_name);
        break;
    default:
        break;
    }
}

int nlk_opts_parse(const struct nlk_opt *opts, const char *usage_line, const char *descr,
                   int argc, const char **argv)
{
    int nrest = 0;
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (a[0] != '-' || a[1] == 0) {     /* not an option */
            argv[nrest++] = a;
            continue;
        }
        if (a[1] == '-' && a[2] == 0) {     /* "--": the rest are arguments */
            for (++i; i < argc; ++i) argv[nrest++] = argv[i];
            break;
        }
        const int is_long = a[1] == '-';
        if ((is_long && !strcmp(a + 2, "help")) || (!is_long && a[1] == 'h' && a[2] == 0)) {
            usage(opts, usage_line, descr);
            exit(0);
        }
        const struct nlk_opt *hit = NULL;
        const char *inline_val = NULL;
        for (const struct nlk_opt *o = opts; o->type != NLK_OPT_END && !hit; ++o) {
            if (o->type == NLK_OPT_GROUP) continue;
            if (is_long && o->long_name) {
                const size_t n = strlen(o->long_name);
                if (!strncmp(a + 2, o->long_name, n) && (a[2 + n] == 0 || a[2 + n] == '=')) {
                    hit = o;
                    if (a[2 + n] == '=') inline_val = a + 3 + n;
                }
            } else if (!is_long && o->short_name && a[1] == o->short_name) {
                hit = o;
                if (a[2]) inline_val = a + 2;
            }
        }
        if (!hit) {
            fprintf(stderr, "error: unknown option `%s`\n", a);
            usage(opts, usage_line, descr);
            exit(1);
        }
        if (!inline_val) {
            if (i + 1 >= argc) die("requires a value", hit->long_name, is_long, hit->short_name);
            inline_val = argv[++i];
        }
        set_value(hit, inline_val, is_long);
    }
    return nrest;
}

int nlk_pick_device(void)
{
    const char *d = getenv("NLK_DEVICE");
    if (!getenv("CUDA_VISIBLE_DEVICES")) {
        setenv("CUDA_VISIBLE_DEVICES", (d && *d) ? d : "0", 1);
        return 0;
    }
    return (d && *d) ? atoi(d) : 0;
}
