/* shmr_map drop-in (src/shmr_map.c): reads-to-contig SHIMMER-pair hits on the GPU through libpgb200.so. */
#include "../include/pgb200.h"
int main(int argc, char **argv) { return pgb_shmr_map_main(argc, argv); }
