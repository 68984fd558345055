"""Worker process hosting ONE build of the unmodified reference extension (module `main`,
src/main.cpp:11-61) from oracle/_ref/<variant>/.  TEST INFRASTRUCTURE ONLY; driven by oracle/ref.py.
The four builds all export the init symbol of a module called `main`, hence one process per variant."""
import pickle
import struct
import sys
import time

import numpy as np


def _read(f):
    hdr = f.read(8)
    if len(hdr) < 8:
        return None
    (n,) = struct.unpack("<Q", hdr)
    return pickle.loads(f.read(n))


def _write(f, obj):
    b = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    f.write(struct.pack("<Q", len(b)))
    f.write(b)
    f.flush()


def serve(variant_dir):
    fin, fout = sys.stdin.buffer, sys.stdout.buffer
    sys.stdout = sys.stderr  # the reference prints to stdout when verbose; keep the pipe clean
    sys.path.insert(0, variant_dir)
    import main  # noqa: the reference extension
    e = None
    empty = np.array([], np.int64)
    while True:
        msg = _read(fin)
        if msg is None:
            return
        op, a = msg
        try:
            if op == "create":
                e = main.RiiCpp(np.ascontiguousarray(a["codewords"], np.float32), False)
                r = main.__version__
            elif op == "add_codes":
                e.add_codes(np.ascontiguousarray(a["codes"], np.uint8), bool(a["update"]))
                r = e.N
            elif op == "reconfigure":
                t0 = time.perf_counter()
                e.reconfigure(int(a["nlist"]), int(a["iter"]))
                r = time.perf_counter() - t0
            elif op == "state":
                r = dict(N=e.N, nlist=e.nlist, coarse_centers=np.array(e.coarse_centers, np.uint8),
                         posting_lists=[np.array(p, np.int32) for p in e.posting_lists])
            elif op == "set_state":  # via the reference's own pickle protocol, src/main.cpp:35-54
                st = (a["codewords"].tolist(), False, a["coarse_centers"].tolist(),
                      a["codes"].reshape(-1).tolist(), [list(map(int, p)) for p in a["posting_lists"]])
                e = main.RiiCpp.__new__(main.RiiCpp)
                e.__setstate__(st)
                r = e.N
            elif op == "query_linear":
                tids = a.get("tids")
                r = e.query_linear(a["q"], int(a["topk"]), empty if tids is None else tids)
            elif op == "query_ivf":
                tids = a.get("tids")
                r = e.query_ivf(a["q"], int(a["topk"]), empty if tids is None else tids, int(a["L"]))
            elif op == "time_queries":  # loop of single-query calls, examples/benchmark/run_sift1m.py:26-30
                Q, topk, L, method = a["Q"], int(a["topk"]), int(a.get("L", 0)), a["method"]
                tids = a.get("tids")
                tids = empty if tids is None else tids
                for q in Q[: int(a.get("warmup", 3))]:
                    e.query_linear(q, topk, tids) if method == "linear" else e.query_ivf(q, topk, tids, L)
                out = []
                t0 = time.perf_counter()
                for q in Q:
                    out.append(e.query_linear(q, topk, tids) if method == "linear"
                               else e.query_ivf(q, topk, tids, L))
                dt = time.perf_counter() - t0
                r = dict(seconds=dt, n=len(Q), ids=[o[0] for o in out] if a.get("return_ids") else None)
            elif op == "time_queries_forked":
                # throughput of `nproc` processes, each a loop of single-query calls over its slice of Q.
                # Children are forked from this process (index pages shared copy-on-write); QueryIvf uses no
                # OpenMP (src/rii.h:261), so forking after the OMP-parallel build is safe for this call.
                import os
                Q, topk, L, nproc = a["Q"], int(a["topk"]), int(a["L"]), int(a["nproc"])
                for q in Q[:3]:
                    e.query_ivf(q, topk, empty, L)
                slices = np.array_split(np.arange(len(Q)), nproc)
                pids, t0 = [], time.perf_counter()
                for sl in slices:
                    pid = os.fork()
                    if pid == 0:
                        try:
                            for i in sl:
                                e.query_ivf(Q[i], topk, empty, L)
                        finally:
                            os._exit(0)
                    pids.append(pid)
                for pid in pids:
                    os.waitpid(pid, 0)
                r = dict(seconds=time.perf_counter() - t0, n=len(Q), nproc=nproc)
            elif op == "quit":
                _write(fout, ("ok", None))
                return
            else:
                raise ValueError(op)
            _write(fout, ("ok", r))
        except BaseException as ex:  # noqa
            _write(fout, ("err", repr(ex)))


if __name__ == "__main__":
    serve(sys.argv[1])
