"""sibeliaz_b200 -- B200-native sibeliaz-lcb hot path (JunctionStorage + BlocksFinder).

Thin ctypes mirror of the reference's C++ call sequence (SibeliaZ-LCB/sibeliaz.cpp:125-143) on top of
the C ABI in include/sibeliaz_lcb.h:

    storage = JunctionStorage(graph, fastas, k, abundance)        # junctionstorage.h:653
    finder = BlocksFinder(storage, k)                             # blocksfinder.h:213
    finder.find_blocks(min_block, max_branch, max_flank)          # blocksfinder.h:453
    finder.generate_output(out_dir, gen_seq, chunks)              # blocksfinder.h:605

Python is plumbing for tests and bench only; all work happens in libsibeliaz_lcb.so (C++/CUDA, sm_100a).
There is no CPU fallback: constructing a BlocksFinder without a usable B200-class GPU raises LcbError.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsibeliaz_lcb.so")
CLI_PATH = os.path.join(_HERE, "bin", "sibeliaz-lcb")

LCB_OK = 0
ERR_NAMES = {1: "LCB_ERR_ARG", 2: "LCB_ERR_IO", 3: "LCB_ERR_FORMAT", 4: "LCB_ERR_CUDA", 5: "LCB_ERR_CAPACITY",
             6: "LCB_ERR_STATE"}

EXPORTS = ["lcb_warmup", "lcb_trim_cache", "lcb_reset_seeds", "lcb_index_load", "lcb_index_get_view", "lcb_index_num_chr", "lcb_index_chr_name", "lcb_index_chr_length",
           "lcb_index_free", "lcb_default_params", "lcb_create", "lcb_comm_unique_id", "lcb_comm_init",
           "lcb_enumerate_seeds", "lcb_get_seeds", "lcb_find_blocks", "lcb_free_blocks", "lcb_get_stats",
           "lcb_last_error", "lcb_destroy", "lcb_write_output", "lcb_version"]


class LcbError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s: %s" % (ERR_NAMES.get(code, code), message))
        self.code = code


class IndexView(C.Structure):
    _fields_ = [("n_chr", C.c_int32), ("n_records", C.c_int64), ("n_vertices", C.c_int64),
                ("chr_off", C.POINTER(C.c_int64)), ("pos_id", C.POINTER(C.c_int32)), ("pos_bp", C.POINTER(C.c_uint32)),
                ("next_ch", C.POINTER(C.c_uint8)), ("prev_rc", C.POINTER(C.c_uint8)), ("vtx_off", C.POINTER(C.c_int64)),
                ("occ_g", C.POINTER(C.c_int64)), ("packed_rec", C.c_void_p), ("packed_occ", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("k", "max_branch", "min_block", "max_flank", "looking_depth", "phase_size",
                                         "window_init", "window_max", "device", "collect_counters")]


class BlockInstance(C.Structure):
    _fields_ = [("id", C.c_int32), ("chr", C.c_uint32), ("start", C.c_uint32), ("end", C.c_uint32)]


BLOCK_DTYPE = np.dtype([("id", "<i4"), ("chr", "<u4"), ("start", "<u4"), ("end", "<u4")])


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_records", "n_vertices", "n_seeds", "n_block_instances", "n_blocks", "windows",
                                          "rounds", "traversals_first", "traversals_rerun", "kernel_launches", "t_walk",
                                          "t_occ", "t_scan", "t_score")] + \
               [("ms_enumerate", C.c_double), ("ms_find", C.c_double), ("ms_traverse_kernels", C.c_double),
                ("traverse_launches", C.c_uint64), ("ms_h2d", C.c_double), ("ms_d2h", C.c_double), ("ms_step_device", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("pool_restarts", C.c_uint64),
                ("big_arena_runs", C.c_uint64), ("lean_runs", C.c_uint64), ("lean_bails", C.c_uint64), ("lean_bail_why", C.c_uint64 * 8), ("ms_tail", C.c_double * 6)]
    # keep in sync with lcb_stats in include/sibeliaz_lcb.h (ms_step_device sits right after ms_d2h)

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n in ("lean_bail_why", "ms_tail") else getattr(self, n)) for n, _ in self._fields_}


class GraphStats(C.Structure):
    # keep in sync with lcg_stats in include/sibeliaz_graph.h
    _fields_ = [(n, C.c_uint64) for n in ("n_records", "n_bases", "n_kmers", "n_distinct", "n_candidates", "n_bifurcations",
                                          "n_junctions", "table_slots", "kernel_launches")] + \
               [(n, C.c_double) for n in ("ms_parse", "ms_h2d", "ms_device", "ms_edges", "ms_d2h", "ms_total")]

    def as_dict(self):
        return {n: (list(getattr(self, n)) if n in ("lean_bail_why", "ms_tail") else getattr(self, n)) for n, _ in self._fields_}


_lib = None


def load_library(path=None):
    """dlopen the in-tree C-ABI library (built by sibeliaz_b200.build); never falls back to anything else."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("LCB_LIB_PATH") or LIB_PATH  # LCB_LIB_PATH: developer builds (build.py --selfcheck)
    if not os.path.exists(path):
        raise LcbError(4, "%s is missing: run `python -m sibeliaz_b200.build` (there is no fallback path)" % path)
    lib = C.CDLL(path)
    lib.lcb_index_load.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    lib.lcb_index_get_view.argtypes = [C.c_void_p, C.POINTER(IndexView)]
    lib.lcb_index_num_chr.argtypes = [C.c_void_p]
    lib.lcb_index_num_chr.restype = C.c_int32
    lib.lcb_index_chr_name.argtypes = [C.c_void_p, C.c_int32]
    lib.lcb_index_chr_name.restype = C.c_char_p
    lib.lcb_index_chr_length.argtypes = [C.c_void_p, C.c_int32]
    lib.lcb_index_chr_length.restype = C.c_int64
    lib.lcb_index_free.argtypes = [C.c_void_p]
    lib.lcb_index_free.restype = None
    lib.lcb_default_params.argtypes = [C.POINTER(Params)]
    lib.lcb_default_params.restype = None
    lib.lcb_create.argtypes = [C.POINTER(IndexView), C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.lcb_comm_unique_id.argtypes = [C.c_void_p]
    lib.lcb_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.lcb_create_shared.argtypes = [C.POINTER(IndexView), C.POINTER(Params), C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.lcb_enumerate_seeds.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    lib.lcb_get_seeds.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    lib.lcb_reset_seeds.argtypes = [C.c_void_p]
    lib.lcb_find_blocks.argtypes = [C.c_void_p, C.POINTER(C.POINTER(BlockInstance)), C.POINTER(C.c_uint64), C.POINTER(Stats)]
    lib.lcb_free_blocks.argtypes = [C.POINTER(BlockInstance)]
    lib.lcb_free_blocks.restype = None
    lib.lcb_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    lib.lcb_last_error.argtypes = [C.c_void_p]
    lib.lcb_last_error.restype = C.c_char_p
    lib.lcb_destroy.argtypes = [C.c_void_p]
    lib.lcb_destroy.restype = None
    lib.lcb_write_output.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                     C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_char_p, C.c_size_t]
    lib.lcb_version.restype = C.c_char_p
    lib.lcg_build_from_fasta.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_void_p),
                                         C.c_char_p, C.c_size_t]
    lib.lcg_build.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_uint64, C.c_int,
                              C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    lib.lcg_build_resident.argtypes = lib.lcg_build.argtypes
    lib.lcb_index_load_fasta.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    lib.lcb_index_get_sequences.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_void_p)), C.POINTER(C.POINTER(C.c_uint64))]
    lib.lcb_index_get_sequences.restype = C.c_int32
    lib.lcb_create_from_graph.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.lcg_num_junctions.argtypes = [C.c_void_p]
    lib.lcg_num_junctions.restype = C.c_uint64
    lib.lcg_get_junctions.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lcg_write_junction_file.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_size_t]
    lib.lcg_get_stats.argtypes = [C.c_void_p, C.POINTER(GraphStats)]
    lib.lcg_free.argtypes = [C.c_void_p]
    lib.lcg_free.restype = None
    if path in (LIB_PATH, os.environ.get("LCB_LIB_PATH")):
        _lib = lib
    return lib


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class JunctionStorage:
    """Host-side junction index (reference: Sibelia::JunctionStorage, junctionstorage.h:116-698)."""

    def __init__(self, graph_file, fasta_files, k, abundance=150):
        lib = load_library()
        self._lib = lib
        self.k = int(k)
        self._h = C.c_void_p()
        files = (C.c_char_p * len(fasta_files))(*[os.fsencode(f) for f in fasta_files])
        err = C.create_string_buffer(1024)
        rc = lib.lcb_index_load(os.fsencode(graph_file), files, len(fasta_files), int(k), int(abundance),
                                C.byref(self._h), err, len(err))
        if rc:
            raise LcbError(rc, err.value.decode(errors="replace"))
        self.view = IndexView()
        lib.lcb_index_get_view(self._h, C.byref(self.view))

    # names follow the reference's accessors
    def get_chr_number(self):
        return self._lib.lcb_index_num_chr(self._h)

    def get_chr_description(self, c):
        return self._lib.lcb_index_chr_name(self._h, c).decode()

    def get_chr_length(self, c):
        return self._lib.lcb_index_chr_length(self._h, c)

    @property
    def n_records(self):
        return self.view.n_records

    @property
    def n_vertices(self):
        return self.view.n_vertices

    def arrays(self):
        """Copies of the SoA arrays as numpy (parity tests)."""
        v = self.view
        N, V, Cn = v.n_records, v.n_vertices, v.n_chr
        mk = lambda p, n, dt: np.ctypeslib.as_array(p, shape=(max(n, 1),))[:n].astype(dt, copy=True) if n else np.zeros(0, dt)
        return dict(chr_off=mk(v.chr_off, Cn + 1, np.int64), pos_id=mk(v.pos_id, N, np.int32), pos_bp=mk(v.pos_bp, N, np.uint32),
                    next_ch=mk(v.next_ch, N, np.uint8), prev_rc=mk(v.prev_rc, N, np.uint8),
                    vtx_off=mk(v.vtx_off, V + 1, np.int64), occ_g=mk(v.occ_g, N, np.int64))

    def close(self):
        if self._h:
            self._lib.lcb_index_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ArrayStorage:
    """An index given directly as numpy SoA arrays (what a foreign host would pass through the C ABI)."""

    def __init__(self, arrays, k):
        self.k = int(k)
        self._keep = {n: np.ascontiguousarray(arrays[n], dt) for n, dt in
                      (("chr_off", np.int64), ("pos_id", np.int32), ("pos_bp", np.uint32), ("next_ch", np.uint8),
                       ("prev_rc", np.uint8), ("vtx_off", np.int64), ("occ_g", np.int64))}
        a = self._keep
        v = IndexView()
        v.n_chr = len(a["chr_off"]) - 1
        v.n_records = len(a["pos_id"])
        v.n_vertices = len(a["vtx_off"]) - 1
        v.chr_off = _ptr(a["chr_off"], C.c_int64)
        v.pos_id = _ptr(a["pos_id"], C.c_int32)
        v.pos_bp = _ptr(a["pos_bp"], C.c_uint32)
        v.next_ch = _ptr(a["next_ch"], C.c_uint8)
        v.prev_rc = _ptr(a["prev_rc"], C.c_uint8)
        v.vtx_off = _ptr(a["vtx_off"], C.c_int64)
        v.occ_g = _ptr(a["occ_g"], C.c_int64)
        self.view = v
        self._h = None

    @property
    def n_records(self):
        return self.view.n_records


class BlocksFinder:
    """Device context (reference: Sibelia::BlocksFinder, blocksfinder.h:178-929)."""

    def __init__(self, storage, k=None, device=0, window_init=0, window_max=0, collect_counters=False, shared=None):
        """`shared` = (rank, n_ranks, id_bytes): multi-GPU creation through lcb_create_shared -- rank 0 uploads the index once,
        the other ranks receive it over NVLink (their `storage` is only consulted for k and may be rank 0's)."""
        self._lib = load_library()
        self._shared = shared
        self.storage = storage
        self.k = int(k if k is not None else storage.k)
        self.device = device
        self._window = (window_init, window_max)
        self._collect = collect_counters
        self._ctx = C.c_void_p()
        self._params = None
        self.blocks = None
        self.stats = None

    def _create(self, min_block, max_branch, max_flank, looking_depth):
        if self._ctx:
            self._lib.lcb_destroy(self._ctx)
            self._ctx = C.c_void_p()
        p = Params()
        self._lib.lcb_default_params(C.byref(p))
        p.k, p.min_block, p.max_branch, p.max_flank, p.looking_depth = self.k, min_block, max_branch, max_flank, looking_depth
        p.device = self.device
        p.window_init, p.window_max = self._window
        p.collect_counters = int(self._collect) if not isinstance(self._collect, bool) else (1 if self._collect else 0)
        if isinstance(self.storage, FusedStorage):
            rc = self._lib.lcb_create_from_graph(self.storage._graph, self.storage._h, self.storage.abundance, C.byref(p), C.byref(self._ctx))
        elif self._shared is not None:
            rank, n_ranks, id_bytes = self._shared
            view = C.byref(self.storage.view) if (rank == 0 or self.storage is not None) else None
            rc = self._lib.lcb_create_shared(view, C.byref(p), rank, n_ranks, id_bytes, C.byref(self._ctx))
        else:
            rc = self._lib.lcb_create(C.byref(self.storage.view), C.byref(p), C.byref(self._ctx))
        if rc:
            msg = self._lib.lcb_last_error(self._ctx).decode() if self._ctx else "lcb_create failed"
            if self._ctx:
                self._lib.lcb_destroy(self._ctx)
                self._ctx = C.c_void_p()
            raise LcbError(rc, msg)
        self._params = p

    def _check(self, rc):
        if rc:
            raise LcbError(rc, self._lib.lcb_last_error(self._ctx).decode())

    def comm_init(self, rank, n_ranks, id_bytes):
        self._check(self._lib.lcb_comm_init(self._ctx, rank, n_ranks, id_bytes))

    def create(self, min_block=200, max_branch=200, max_flank=None, looking_depth=8):
        self._create(int(min_block), int(max_branch), int(max_branch if max_flank is None else max_flank), int(looking_depth))
        return self

    def enumerate_seeds(self):
        if not self._ctx:
            self.create()
        n = C.c_uint64()
        self._check(self._lib.lcb_enumerate_seeds(self._ctx, C.byref(n)))
        return n.value

    def seeds(self):
        n = self.enumerate_seeds()
        out = dict(vid=np.zeros(n, np.int64), ch=np.zeros(n, np.uint8), count=np.zeros(n, np.uint64),
                   rank=np.zeros(n, np.uint64), res_pos=np.zeros(n, np.uint64), res_chr=np.zeros(n, np.uint64))
        self._check(self._lib.lcb_get_seeds(self._ctx, *[out[k].ctypes.data for k in ("vid", "ch", "count", "rank", "res_pos", "res_chr")]))
        return out

    def find_blocks(self, min_block=200, max_branch=200, max_flank=None, looking_depth=8, sample_size=0, threads=1,
                    debug_out=""):
        """FindBlocks(minBlockSize, maxBranchSize, maxFlankingSize, lookingDepth, sampleSize, threads, debugOut);
        the last three are accepted for signature parity and ignored (the reference ignores sampleSize/debugOut too)."""
        if not self._ctx or self._params is None or (self._params.min_block, self._params.max_branch) != (int(min_block), int(max_branch)):
            self.create(min_block, max_branch, max_flank, looking_depth)
        ptr = C.POINTER(BlockInstance)()
        n = C.c_uint64()
        st = Stats()
        self._check(self._lib.lcb_find_blocks(self._ctx, C.byref(ptr), C.byref(n), C.byref(st)))
        if n.value:
            self.blocks = np.empty(n.value, BLOCK_DTYPE)  # one copy out of the library's (page-locked) buffer
            C.memmove(self.blocks.ctypes.data, ptr, n.value * C.sizeof(BlockInstance))
        else:
            self.blocks = np.zeros(0, BLOCK_DTYPE)
        self._lib.lcb_free_blocks(ptr)
        self.stats = st.as_dict()
        return self.blocks

    def reset_seeds(self):
        self._check(self._lib.lcb_reset_seeds(self._ctx))

    def generate_output(self, out_dir, gen_seq=False, chunks=0, min_block=None):
        if self.blocks is None:
            raise LcbError(6, "find_blocks has not run")
        if self.storage._h is None:
            raise LcbError(6, "generate_output needs a JunctionStorage loaded from files")
        found, cov = C.c_int64(), C.c_double()
        err = C.create_string_buffer(1024)
        b = np.ascontiguousarray(self.blocks)
        m = self._params.min_block if min_block is None else int(min_block)
        rc = self._lib.lcb_write_output(self.storage._h, b.ctypes.data, len(b), m, os.fsencode(out_dir), int(bool(gen_seq)),
                                        int(chunks), C.byref(found), C.byref(cov), err, len(err))
        if rc:
            raise LcbError(rc, err.value.decode(errors="replace"))
        return found.value, cov.value

    def close(self):
        if self._ctx:
            self._lib.lcb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class JunctionGraph:
    """Junctions of the compacted de Bruijn graph of a set of FASTA files, found on the GPU (reference: the `twopaco`
    step of the pipeline, TwoPaCo VertexEnumeratorImpl, vertexenumerator.h:122-466).  `sequences=` takes in-memory
    records (bytes) instead of files."""

    def __init__(self, fastas=None, k=25, device=0, abundance=2 ** 64 - 1, sequences=None):
        self._lib = load_library()
        self._h = C.c_void_p()
        err = C.create_string_buffer(1024)
        if sequences is not None:
            bufs = [np.frombuffer(bytes(s), dtype=np.uint8) for s in sequences]
            ptrs = (C.c_void_p * len(bufs))(*[b.ctypes.data if len(b) else None for b in bufs])
            lens = (C.c_uint64 * len(bufs))(*[len(b) for b in bufs])
            rc = self._lib.lcg_build(ptrs, lens, len(bufs), int(k), int(abundance), int(device), C.byref(self._h), err, len(err))
        else:
            files = (C.c_char_p * len(fastas))(*[os.fsencode(f) for f in fastas])
            rc = self._lib.lcg_build_from_fasta(files, len(fastas), int(k), int(abundance), int(device), C.byref(self._h), err, len(err))
        if rc:
            raise LcbError(rc, err.value.decode(errors="replace"))
        st = GraphStats()
        self._lib.lcg_get_stats(self._h, C.byref(st))
        self.stats = st.as_dict()
        self.k = int(k)

    def junctions(self):
        n = self._lib.lcg_num_junctions(self._h)
        out = dict(chr=np.zeros(n, np.uint32), pos=np.zeros(n, np.uint32), id=np.zeros(n, np.int64))
        self._lib.lcg_get_junctions(self._h, out["chr"].ctypes.data, out["pos"].ctypes.data, out["id"].ctypes.data)
        return out

    def write(self, path):
        err = C.create_string_buffer(1024)
        rc = self._lib.lcg_write_junction_file(self._h, os.fsencode(path), err, len(err))
        if rc:
            raise LcbError(rc, err.value.decode(errors="replace"))
        return path

    def close(self):
        if self._h:
            self._lib.lcg_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


GRAPH_CLI_PATH = os.path.join(_HERE, "bin", "twopaco")
ALIGN_CLI_PATH = os.path.join(_HERE, "bin", "sibeliaz-align")


class FusedStorage:
    """The fused pipeline's stand-in for JunctionStorage: FASTA files only.  The junctions are found on the GPU
    (lcg_build_resident) and stay there; BlocksFinder builds the junction index on the device from them
    (lcb_create_from_graph) -- no junction file, no host-side index."""

    def __init__(self, fastas, k, abundance=150, device=0):
        self._lib = load_library()
        self.k, self.abundance, self.device = int(k), int(abundance), int(device)
        self._h, self._graph = C.c_void_p(), C.c_void_p()
        err = C.create_string_buffer(1024)
        files = (C.c_char_p * len(fastas))(*[os.fsencode(f) for f in fastas])
        rc = self._lib.lcb_index_load_fasta(files, len(fastas), self.k, C.byref(self._h), err, len(err))
        if rc:
            raise LcbError(rc, err.value.decode(errors="replace"))
        seq, ln = C.POINTER(C.c_void_p)(), C.POINTER(C.c_uint64)()
        n = self._lib.lcb_index_get_sequences(self._h, C.byref(seq), C.byref(ln))
        rc = self._lib.lcg_build_resident(seq, ln, n, self.k, 2 ** 64 - 1, self.device, C.byref(self._graph), err, len(err))
        if rc:
            self.close()
            raise LcbError(rc, err.value.decode(errors="replace"))
        st = GraphStats()
        self._lib.lcg_get_stats(self._graph, C.byref(st))
        self.graph_stats = st.as_dict()

    def close(self):
        if self._graph:
            self._lib.lcg_free(self._graph)
            self._graph = C.c_void_p()
        if self._h:
            self._lib.lcb_index_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------------------
# Alignment stage (include/sibeliaz_align.h): `spoa` per block -> alignment.maf, on the GPU
class AlignParams(C.Structure):
    _fields_ = [("match", C.c_int), ("mismatch", C.c_int), ("gap", C.c_int), ("device", C.c_int)]


class AlignStats(C.Structure):
    # keep in sync with lca_stats in include/sibeliaz_align.h
    _fields_ = [("n_blocks", C.c_uint64), ("n_copies", C.c_uint64), ("n_bases", C.c_uint64), ("cells", C.c_uint64),
                ("blocks_level", C.c_uint64 * 3), ("kernel_launches", C.c_uint64), ("ms_kernels", C.c_double),
                ("ms_total", C.c_double), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "blocks_level"}
        d["blocks_level"] = list(self.blocks_level)
        return d


def _align_lib():
    lib = load_library()
    if not getattr(lib, "_lca_ready", False):
        lib.lca_default_params.argtypes = [C.POINTER(AlignParams)]
        lib.lca_default_params.restype = None
        lib.lca_align.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.POINTER(AlignParams),
                                  C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
        lib.lca_rows.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint64))]
        lib.lca_rows.restype = C.POINTER(C.c_uint8)
        lib.lca_block_columns.argtypes = [C.c_void_p, C.c_uint32]
        lib.lca_block_columns.restype = C.c_uint32
        lib.lca_get_stats.argtypes = [C.c_void_p, C.POINTER(AlignStats)]
        lib.lca_get_stats.restype = None
        lib.lca_free.argtypes = [C.c_void_p]
        lib.lca_free.restype = None
        lib.lca_align_chunk_files.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_char_p, C.POINTER(AlignParams),
                                              C.POINTER(AlignStats), C.c_char_p, C.c_size_t]
        lib._lca_ready = True
    return lib


def align_blocks(blocks, match=5, mismatch=-4, gap=-8, device=0):
    """MSA rows of every block (reference: `spoa <block.fa> -l 1 -r 1 -e -8`, one run per block).
    blocks: list of lists of bytes (the copies of a block in file order).  Returns (rows, stats): rows[b][k] = bytes."""
    lib = _align_lib()
    copies = [c for b in blocks for c in b]
    seq = np.frombuffer(b"".join(copies), dtype=np.uint8) if copies else np.zeros(0, np.uint8)
    copy_off = np.zeros(len(copies) + 1, np.uint64)
    if copies:
        copy_off[1:] = np.cumsum([len(c) for c in copies])
    block_off = np.zeros(len(blocks) + 1, np.uint32)
    if blocks:
        block_off[1:] = np.cumsum([len(b) for b in blocks])
    p = AlignParams()
    lib.lca_default_params(C.byref(p))
    p.match, p.mismatch, p.gap, p.device = match, mismatch, gap, device
    res, err = C.c_void_p(), C.create_string_buffer(1024)
    rc = lib.lca_align(seq.ctypes.data, copy_off.ctypes.data, len(copies), block_off.ctypes.data, len(blocks), C.byref(p), C.byref(res),
                       err, len(err))
    if rc:
        raise LcbError(rc, err.value.decode(errors="replace"))
    try:
        roff = C.POINTER(C.c_uint64)()
        rows = lib.lca_rows(res, C.byref(roff))
        off = np.ctypeslib.as_array(roff, shape=(len(copies) + 1,)).copy() if copies else np.zeros(1, np.uint64)
        total = int(off[-1])
        flat = bytes(np.ctypeslib.as_array(rows, shape=(total,))) if total else b""
        out, c = [], 0
        for b in blocks:
            out.append([flat[int(off[c + k]):int(off[c + k + 1])] for k in range(len(b))])
            c += len(b)
        st = AlignStats()
        lib.lca_get_stats(res, C.byref(st))
        return out, st.as_dict()
    finally:
        lib.lca_free(res)


def global_alignment(chunk_files, cmd, out_maf, match=5, mismatch=-4, gap=-8, device=0):
    """The wrapper's global_alignment() (SibeliaZ-LCB/sibeliaz:118-134): <i>.tmp chunk files -> alignment.maf."""
    lib = _align_lib()
    p = AlignParams()
    lib.lca_default_params(C.byref(p))
    p.match, p.mismatch, p.gap, p.device = match, mismatch, gap, device
    files = (C.c_char_p * len(chunk_files))(*[os.fsencode(f) for f in chunk_files])
    st, err = AlignStats(), C.create_string_buffer(1024)
    rc = lib.lca_align_chunk_files(files, len(chunk_files), cmd.encode(), os.fsencode(out_maf), C.byref(p), C.byref(st), err, len(err))
    if rc:
        raise LcbError(rc, err.value.decode(errors="replace"))
    return st.as_dict()
