"""ctypes binding of the CPU restatement oracle (oracle/liblcb_oracle.so) and helpers around the compiled
reference (oracle/_ref).  TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the sibeliaz_b200 package."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liblcb_oracle.so")
REF_LCB = os.path.join(ORACLE_DIR, "_ref", "sibeliaz-lcb-ref")
REF_TWOPACO = os.path.join(ORACLE_DIR, "_ref", "twopaco")
POA_ORACLE = os.path.join(ORACLE_DIR, "poa_oracle")  # CPU restatement of the alignment stage (oracle/poa_oracle.cpp)
REF_SPOA = os.path.join(ORACLE_DIR, "_ref", "spoa-ref")  # the reference's spoa library behind oracle/spoa_driver.cpp

_lib = None


def build_oracle():
    subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if os.path.isdir("/root/reference/SibeliaZ-LCB"):
        subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            build_oracle()
        L = C.CDLL(ORACLE_LIB)
        L.lcbo_load.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.lcbo_load.restype = C.c_void_p
        L.lcbo_free.argtypes = [C.c_void_p]
        L.lcbo_free.restype = None
        for f in ("lcbo_num_records", "lcbo_num_vertices"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_int64
        L.lcbo_num_chr.argtypes = [C.c_void_p]
        L.lcbo_num_chr.restype = C.c_int32
        L.lcbo_get_index.argtypes = [C.c_void_p] + [C.c_void_p] * 8
        L.lcbo_get_index.restype = None
        L.lcbo_enumerate_seeds.argtypes = [C.c_void_p]
        L.lcbo_enumerate_seeds.restype = C.c_int64
        L.lcbo_get_seeds.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.lcbo_get_seeds.restype = None
        L.lcbo_find_blocks.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.lcbo_find_blocks.restype = C.c_int64
        L.lcbo_get_blocks.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.lcbo_get_blocks.restype = None
        L.lcbo_get_counters.argtypes = [C.c_void_p, C.c_void_p]
        L.lcbo_get_counters.restype = None
        L.lcbo_generate_output.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64),
                                           C.POINTER(C.c_double), C.c_char_p, C.c_int]
        _lib = L
    return _lib


class Oracle:
    def __init__(self, graph, fastas, k, abundance=150):
        L = lib()
        err = C.create_string_buffer(512)
        files = (C.c_char_p * len(fastas))(*[os.fsencode(f) for f in fastas])
        self.h = L.lcbo_load(os.fsencode(graph), files, len(fastas), k, abundance, err, len(err))
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.k = k
        self.N, self.V, self.C = L.lcbo_num_records(self.h), L.lcbo_num_vertices(self.h), L.lcbo_num_chr(self.h)

    def index(self):
        a = dict(chr_off=np.zeros(self.C + 1, np.int64), pos_id=np.zeros(self.N, np.int32), pos_bp=np.zeros(self.N, np.uint32),
                 next_ch=np.zeros(self.N, np.uint8), prev_rc=np.zeros(self.N, np.uint8), vtx_off=np.zeros(self.V + 1, np.int64),
                 occ_g=np.zeros(self.N, np.int64), chr_len=np.zeros(self.C, np.int64))
        lib().lcbo_get_index(self.h, *[a[k].ctypes.data for k in ("chr_off", "pos_id", "pos_bp", "next_ch", "prev_rc", "vtx_off", "occ_g", "chr_len")])
        return a

    def seeds(self):
        n = lib().lcbo_enumerate_seeds(self.h)
        out = dict(vid=np.zeros(n, np.int64), ch=np.zeros(n, np.uint8), count=np.zeros(n, np.uint64), rank=np.zeros(n, np.uint64),
                   res_pos=np.zeros(n, np.uint64), res_chr=np.zeros(n, np.uint64))
        lib().lcbo_get_seeds(self.h, *[out[k].ctypes.data for k in ("vid", "ch", "count", "rank", "res_pos", "res_chr")])
        return out

    def find_blocks(self, min_block, max_branch, max_flank=None, looking_depth=8, phase=256):
        n = lib().lcbo_find_blocks(self.h, min_block, max_branch, max_branch if max_flank is None else max_flank, looking_depth, phase)
        ids, ch, st, en = np.zeros(n, np.int32), np.zeros(n, np.uint32), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        lib().lcbo_get_blocks(self.h, ids.ctypes.data, ch.ctypes.data, st.ctypes.data, en.ctypes.data)
        ctr = np.zeros(8, np.uint64)
        lib().lcbo_get_counters(self.h, ctr.ctypes.data)
        self.counters = dict(zip(("t_walk", "t_occ", "t_scan", "t_score", "process", "reruns", "mpv", "pushes"), ctr.tolist()))
        return dict(id=ids, chr=ch, start=st, end=en)

    def generate_output(self, outdir, gen_seq, chunks, min_block):
        found, cov = C.c_int64(), C.c_double()
        err = C.create_string_buffer(512)
        rc = lib().lcbo_generate_output(self.h, os.fsencode(outdir), int(gen_seq), chunks, min_block, C.byref(found), C.byref(cov), err, len(err))
        if rc:
            raise RuntimeError(err.value.decode())
        return found.value, cov.value

    def close(self):
        if self.h:
            lib().lcbo_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run_twopaco(fastas, k, out, threads=8, tmpdir=None):
    tmpdir = tmpdir or os.path.dirname(out)
    subprocess.run([REF_TWOPACO, "--tmpdir", tmpdir, "-t", str(threads), "-k", str(k), "--filtermemory", "4", "-o", out] + list(fastas),
                   check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return out


def run_reference_lcb(graph, fastas, k, outdir, b=200, m=50, a=150, threads=1, noseq=True, chunks=0):
    cmd = [REF_LCB, "--graph", graph] + list(fastas) + ["-k", str(k), "-b", str(b), "-o", outdir, "-m", str(m), "-t", str(threads),
                                                        "--abundance", str(a)]
    cmd += ["--noseq"] if noseq else ["--chunks", str(chunks)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError(r.stdout)
    return r.stdout


# ---- graph construction (oracle/graph_oracle.cpp): the step before the LCB path -----------------------------------
GRAPH_ORACLE_LIB = os.path.join(ORACLE_DIR, "libgraph_oracle.so")
_glib = None


def graph_lib():
    global _glib
    if _glib is None:
        if not os.path.exists(GRAPH_ORACLE_LIB):
            build_oracle()
        L = C.CDLL(GRAPH_ORACLE_LIB)
        L.gro_build.argtypes = [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_uint64, C.c_char_p, C.c_char_p, C.c_int]
        L.gro_build.restype = C.c_int64
        L.gro_canonicalize.argtypes = [C.c_char_p, C.c_char_p]
        L.gro_canonicalize.restype = C.c_int64
        _glib = L
    return _glib


def graph_oracle_build(fastas, k, out, abundance=2 ** 64 - 1):
    """Junction file of the FASTA files by the CPU restatement; returns the number of records."""
    err = C.create_string_buffer(512)
    files = (C.c_char_p * len(fastas))(*[os.fsencode(f) for f in fastas])
    n = graph_lib().gro_build(files, len(fastas), k, abundance, os.fsencode(out), err, len(err))
    if n < 0:
        raise RuntimeError(err.value.decode())
    return n


def canonical_junctions(path, out=None):
    """Label-free normal form of a junction file (vertices renumbered by first appearance, first appearance positive):
    returns its bytes.  Two junction files describe the same graph iff these are equal."""
    out = out or path + ".canon"
    if graph_lib().gro_canonicalize(os.fsencode(path), os.fsencode(out)) < 0:
        raise RuntimeError("cannot canonicalise " + path)
    with open(out, "rb") as f:
        return f.read()


def maf_paragraphs(path_or_text, is_text=False):
    """{key: paragraph} of a MAF file; key = the (name, start, length, strand) tuples of its rows, paragraph = its `s` lines."""
    text = path_or_text if is_text else open(path_or_text).read()
    out, cur = {}, None
    for line in text.splitlines():
        if line.startswith("a"):
            cur = []
        elif line.startswith("s ") and cur is not None:
            cur.append(line)
        elif cur:
            out[tuple(tuple(x.split(" ")[1:5]) for x in cur)] = cur
            cur = None
    if cur:
        out[tuple(tuple(x.split(" ")[1:5]) for x in cur)] = cur
    return out


def reference_global_alignment(outdir, cmd, chunk_files=None):
    """The wrapper's global_alignment() (SibeliaZ-LCB/sibeliaz:118-134) with the reference's spoa library: every line of
    every <i>.tmp chunk is one block, aligned by `spoa -l 1 -r 1 -e -8`; the per-chunk results are concatenated in the
    C-locale order of the file names behind the three header lines.  Returns the MAF text.  `chunk_files` restricts the
    run to some chunk files (tests)."""
    names = sorted(n for n in os.listdir(outdir) if n.endswith(".tmp")) if chunk_files is None else list(chunk_files)
    text = "##maf version=1\n# sibeliaz v1.2.7 \n# cmd=%s\n" % cmd
    for n in sorted(names):  # Python compares str by code point == LC_ALL=C for these ASCII names
        r = subprocess.run([REF_SPOA, "--chunk", os.path.join(outdir, n), "-l", "1", "-r", "1", "-e", "-8"], check=True,
                           stdout=subprocess.PIPE, text=True)
        text += r.stdout
    return text


def poa_oracle_text(chunk_file):
    """MAF paragraphs of every block of one chunk file, by the CPU restatement (oracle/poa_oracle.cpp)."""
    if not os.path.exists(POA_ORACLE):
        build_oracle()
    return subprocess.run([POA_ORACLE, "--chunk", chunk_file], check=True, stdout=subprocess.PIPE, text=True).stdout


def write_chunk(path, blocks):
    """An LCB chunk file (blocksfinder.h:533-582) from [[(header, sequence), ...], ...]; header e.g. 'g0.chr1;10;4;+;100'."""
    with open(path, "w") as f:
        for b in blocks:
            f.write("".join("> %s@%s@" % (h, s) for h, s in b) + "\n")
    return path
