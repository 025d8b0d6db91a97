/* fixed_time.c -- link-time stand-in for time(): the reference's main() seeds rand() from the
 * clock (qpsk.c:294); interposing a constant makes two builds of the unmodified source comparable. */
#include <time.h>
time_t time(time_t *t) { if (t) *t = (time_t)12345; return (time_t)12345; }
