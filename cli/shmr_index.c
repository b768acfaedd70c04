/* shmr_index — drop-in for the reference tool of the same name (src/shmr_index.c:37-245): same options, defaults,
 * output files and messages; the work runs on the GPU through libpgb200.so (include/pgb200.h). */
#include "../include/pgb200.h"
int main(int argc, char **argv) { return pgb_shmr_index_main(argc, argv); }
