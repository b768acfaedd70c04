"""Drop-in for ``peregrine._shimmer4py``: a cffi ``ffi`` / ``lib`` pair whose ``lib`` is libpgb200.so.

The reference builds its module in API mode from its own C sources (py/peregrine/build_shimmer4py.py:8-96); here the same
C declarations are bound in ABI mode to the CUDA library, so reference-style Python keeps working:

    from peregrine_b200.shimmer4py import ffi, lib
    mmers = ffi.new("mm128_v *")
    lib.mm_sketch(ffi.NULL, seq, len(seq), 80, 16, 0, 0, mmers)        # py/peregrine/utils.py:28-49
    lib.free(mmers.a)
"""
from cffi import FFI

from .engine import lib_path

ffi = FFI()
ffi.cdef("""
typedef int32_t seq_coor_t;
typedef struct { seq_coor_t m_size, dist; seq_coor_t q_bgn, q_end; seq_coor_t t_bgn, t_end; seq_coor_t t_m_end, q_m_end; } ovlp_match_t;
typedef struct { uint64_t x, y; } mm128_t;
typedef struct { size_t n, m; mm128_t *a; } mm128_v;
typedef struct { mm128_v *mmers; void *mmer0_map; void *rlmap; void *mcmap; void *ridmm; } py_mmer_t;
typedef struct { uint64_t x0, x1, y0, y1; uint8_t direction; } mp256_t;
typedef struct { size_t n, m; mp256_t *a; } mp256_v;
typedef uint32_t mm_idx_t;
typedef struct { size_t n, m; mm_idx_t *a; } mm_idx_v;
typedef struct { mm_idx_v idx0; mm_idx_v idx1; } shmr_aln_t;
typedef struct { size_t n, m; shmr_aln_t *a; } shmr_aln_v;

void decode_biseq(uint8_t *src, char *seq, size_t len, uint8_t strand);
ovlp_match_t *ovlp_match(uint8_t *query_seq, seq_coor_t q_len, uint8_t q_strand, uint8_t *target_seq, seq_coor_t t_len,
                         uint8_t t_strand, seq_coor_t band_tolerance);
void free_ovlp_match(ovlp_match_t *match);
mm128_v read_mmlist(char *fn);
void mm_sketch(void *km, const char *str, int len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p);
void mm_reduce(mm128_v *in, mm128_v *out, uint8_t rs);
shmr_aln_v *shmr_aln(mm128_v *, mm128_v *, uint8_t, uint32_t, uint32_t, uint32_t);
void free_shmr_alns(shmr_aln_v *);
void build_shimmer_map4py(py_mmer_t *, char *, char *, uint32_t, uint32_t, uint32_t, uint32_t);
void get_shimmers_for_read(mm128_v *, py_mmer_t *, uint32_t);
uint32_t get_mmer_count(py_mmer_t *, uint64_t);
void get_shimmer_hits(mp256_v *, py_mmer_t *, uint64_t, uint32_t);
void free(void *ptr);
""")


class _Lazy:
    """dlopen on first attribute access, so that importing this module never needs the GPU or the built library."""

    _lib = None

    def __getattr__(self, name):
        if _Lazy._lib is None:
            _Lazy._lib = ffi.dlopen(lib_path())
        return getattr(_Lazy._lib, name)


lib = _Lazy()
