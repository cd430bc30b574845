/* C-only consumer of the drop-in boundary (include/airmodes_b200.h): read a cfile (interleaved float32 I/Q, what
 * `modes_rx -s file` feeds to rx_path, python/radio.py:221-232), push it through amb_process in chunks and print
 * the slicer messages exactly as slicer_impl.cc:186-194 would queue them.
 *
 *   gcc -O2 -Iinclude examples/modes_rx_c.c -Lgr_air_modes_b200 -lairmodes_b200 -Wl,-rpath,$PWD/gr_air_modes_b200 -o modes_rx_c
 *   ./modes_rx_c capture.cfile 4e6 7.0
 */
#include <stdio.h>
#include <stdlib.h>
#include "airmodes_b200.h"

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s file.cfile rate [threshold_db] [chunk_samples]\n", argv[0]); return 2; }
    const float rate = (float)atof(argv[2]);
    const float thr = argc > 3 ? (float)atof(argv[3]) : 7.0f;           /* radio.py:114 */
    const size_t chunk = argc > 4 ? (size_t)atoll(argv[4]) : (size_t)1 << 22;
    amb_ctx* ctx = NULL;
    int rc = amb_create(0, rate, thr, /*use_pmf=*/1, /*use_dcblock=*/0, &ctx);
    if (rc != AMB_OK) { fprintf(stderr, "amb_create: %s\n", amb_strerror(rc)); return 1; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    float* buf = (float*)malloc(chunk * 2 * sizeof(float));
    amb_frame frames[512];
    char text[160];
    int first = 1;                                                       /* slicer_impl.cc:192: precision quirk */
    for (;;) {
        const size_t got = fread(buf, 2 * sizeof(float), chunk, f);
        const int last = got < chunk;
        rc = amb_process(ctx, buf, got, AMB_MEM_HOST, last);
        if (rc != AMB_OK) { fprintf(stderr, "amb_process: %s (%s)\n", amb_strerror(rc), amb_last_error(ctx)); return 1; }
        int n;
        while ((n = amb_poll_frames(ctx, frames, 512)) > 0)
            for (int k = 0; k < n; k++)
                if (frames[k].passed) {
                    amb_format_message(&frames[k], first, text, sizeof text);
                    first = 0;
                    puts(text);
                }
        if (n < 0) { fprintf(stderr, "amb_poll_frames: %s\n", amb_strerror(n)); return 1; }
        if (last) break;
    }
    free(buf);
    fclose(f);
    amb_destroy(ctx);
    return 0;
}
