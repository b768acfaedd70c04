/* shmr_overlap — drop-in for the reference tool of the same name (src/shmr_overlap.c:233-419): same options,
 * defaults, input globbing and output stream; the work runs on the GPU through libpgb200.so (include/pgb200.h). */
#include "../include/pgb200.h"
int main(int argc, char **argv) { return pgb_shmr_overlap_main(argc, argv); }
