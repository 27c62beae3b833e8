/* nlk_opts.h -- command-line options of the host drivers.
 *
 * Keeps the surface the reference drivers get from lib/argparse (reference
 * lib/argparse/argparse.c): "--name value", "--name=value", "-x value", "-xvalue";
 * the LAST occurrence of an option wins (the pipeline scripts rely on it:
 * `$FPM --f2_p 0`, reference scripts/nlkalman-seq.sh:80); "--" ends the options;
 * "-h"/"--help" prints the option table on stdout and exits 0; an unknown option or
 * a malformed value prints "error: ..." on stderr and exits 1.
 */
#ifndef NLK_OPTS_H
#define NLK_OPTS_H

enum nlk_opt_type { NLK_OPT_END = 0, NLK_OPT_GROUP, NLK_OPT_STRING, NLK_OPT_INT, NLK_OPT_FLOAT };

struct nlk_opt {
    enum nlk_opt_type type;
    char short_name;        /* 0 = none */
    const char *long_name;  /* for NLK_OPT_GROUP: the heading */
    void *value;            /* const char **, int * or float * */
    const char *help;
};

/* Parses argv[1..]; returns the number of remaining (non-option) arguments, moved to
 * argv[0..]. */
int nlk_opts_parse(const struct nlk_opt *opts, const char *usage, const char *descr,
                   int argc, const char **argv);

/* The device of this process: NLK_DEVICE (default 0).  Unless the caller has restricted the visible
 * devices already, the others are hidden before the CUDA runtime comes up -- it initialises every
 * visible GPU, which on an 8-GPU box is most of the wall time of a per-frame invocation.  Returns the
 * device index to pass to nlk_ctx_create. */
int nlk_pick_device(void);

#endif
