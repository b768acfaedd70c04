/* shmr_mkseqdb drop-in (src/shmr_mkseqdb.c): -d <file list> -p <prefix> -> <prefix>.idx + <prefix>.seqdb; bases are encoded on
 * the GPU through libpgb200.so. */
#include "../include/pgb200.h"
int main(int argc, char **argv) { return pgb_shmr_mkseqdb_main(argc, argv); }
