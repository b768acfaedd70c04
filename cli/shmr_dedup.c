/* shmr_dedup drop-in: `cat ovlp-*.dat | shmr_dedup > preads.ovl` (py/scripts/pg_run.py:352) on the GPU via libpgb200.so. */
#include "../include/pgb200.h"
int main(int argc, char **argv) { return pgb_shmr_dedup_main(argc, argv); }
