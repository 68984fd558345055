"""Driver for the UNMODIFIED reference (`main.RiiCpp`, src/main.cpp:12-54) compiled into oracle/_ref/ by
oracle/Makefile.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's reference/cpu_baseline legs).

    r = Ref("strict")            # or "fast"; picks the _v4 (AVX-512) or _v3 (AVX2) build for this host
    r.create(codewords); r.add_codes(codes, False); r.reconfigure(nlist, iter)
    ids, dists = r.query_linear(q, topk, tids); r.query_ivf(q, topk, tids, L)

`strict` (-O2 -ffp-contract=off) executes the reference arithmetic as written: the bit-exact semantic
oracle.  `fast` (-Ofast, the reference's own setup.py flags) is what users run: the CPU timing baseline.
"""
import glob
import os
import pickle
import struct
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _cpu_has_avx512():
    try:
        return " avx512f" in open("/proc/cpuinfo").read()
    except OSError:
        return False


def variant_dir(kind):
    """kind: 'strict' | 'fast' (auto ISA) or an explicit 'strict_v3' etc.  Returns None if not built."""
    if "_" not in kind:
        kind = kind + ("_v4" if _cpu_has_avx512() else "_v3")
    d = os.path.join(_HERE, "_ref", kind)
    return d if glob.glob(os.path.join(d, "main*.so")) else None


def available(kind="strict"):
    return variant_dir(kind) is not None


def simd_width(kind):
    """Accumulator width of fvec_L2sqr in that build (src/distance.h:113,172)."""
    d = variant_dir(kind)
    return 16 if d and d.endswith("_v4") else 8


class Ref(object):
    def __init__(self, kind="strict", omp_threads=None):
        d = variant_dir(kind)
        if d is None:
            raise RuntimeError("oracle/_ref/%s is not built (run `make -C oracle ref` where /root/reference exists)" % kind)
        self.kind = os.path.basename(d)
        env = dict(os.environ)
        if omp_threads:
            env["OMP_NUM_THREADS"] = str(omp_threads)
        self.p = subprocess.Popen([sys.executable, os.path.join(_HERE, "_ref_worker.py"), d],
                                  stdin=subprocess.PIPE, stdout=subprocess.PIPE, env=env)

    def _call(self, op, **a):
        b = pickle.dumps((op, a), protocol=pickle.HIGHEST_PROTOCOL)
        self.p.stdin.write(struct.pack("<Q", len(b)))
        self.p.stdin.write(b)
        self.p.stdin.flush()
        hdr = self.p.stdout.read(8)
        if len(hdr) < 8:
            raise RuntimeError("reference worker died (rc=%s)" % self.p.poll())
        (n,) = struct.unpack("<Q", hdr)
        st, r = pickle.loads(self.p.stdout.read(n))
        if st != "ok":
            raise RuntimeError("reference worker: " + str(r))
        return r

    def create(self, codewords):
        return self._call("create", codewords=codewords)

    def add_codes(self, codes, update=False):
        return self._call("add_codes", codes=codes, update=update)

    def reconfigure(self, nlist, iter=5):
        return self._call("reconfigure", nlist=nlist, iter=iter)

    def state(self):
        return self._call("state")

    def set_state(self, codewords, coarse_centers, codes, posting_lists):
        return self._call("set_state", codewords=codewords, coarse_centers=coarse_centers, codes=codes,
                          posting_lists=posting_lists)

    def query_linear(self, q, topk, tids=None):
        return self._call("query_linear", q=q, topk=topk, tids=tids)

    def query_ivf(self, q, topk, tids, L):
        return self._call("query_ivf", q=q, topk=topk, tids=tids, L=L)

    def time_queries(self, Q, topk, method, L=0, tids=None, warmup=3, return_ids=False):
        return self._call("time_queries", Q=Q, topk=topk, method=method, L=L, tids=tids, warmup=warmup,
                          return_ids=return_ids)

    def close(self):
        if self.p and self.p.poll() is None:
            try:
                self._call("quit")
            except Exception:
                pass
            self.p.wait(timeout=10)
        self.p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
