"""ctypes binding of oracle/liboracle.so, our C restatement of the PDict path, exposed
with the same surface as oracle/ref.py's RefPPTB  (TEST INFRASTRUCTURE ONLY).

The per-subject core is oracle/pdict_oracle.c::oracle_match; the drivers for views and
subject sets are restated here from src/match_pdict.c:225-264 (views: each view is an
independent subject, ends shifted by the view offset and appended in view order),
:320-346 (vwhich) and :386-432 with :85-127 (vcount, collapse/weight).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

NA_INTEGER = -2147483648
MATCHES_AS_WHICH, MATCHES_AS_COUNTS, MATCHES_AS_ENDS = 1, 2, 4
DEFAULT_MAX_WORK = 5_000_000     # per-window expansion budget (fixedS=FALSE)

_lib = None


class PortError(RuntimeError):
    pass


def build_lib():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO) or \
                os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "pdict_oracle.c")):
            build_lib()
        L = C.CDLL(_SO)
        L.oracle_pptb_build.restype = C.c_void_p
        L.oracle_pptb_build.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]
        L.oracle_pptb_free.argtypes = [C.c_void_p]
        L.oracle_pptb_blank_dups.argtypes = [C.c_void_p]
        L.oracle_pptb_nnodes.argtypes = [C.c_void_p]
        L.oracle_pptb_nnodes.restype = C.c_long
        L.oracle_match.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_void_p, C.c_int] + \
            [C.c_int] * 4 + [C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_char_p, C.c_int]
        L.oracle_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class PortPPTB:
    def __init__(self, algo, tb_concat, tb_widths, pp_exclude=None):
        self.algo = algo
        tb, tbw = _u8(tb_concat), _i32(tb_widths)
        self.M = len(tbw)
        excl = None if pp_exclude is None else _i32(pp_exclude)
        self.high2low = np.empty(self.M, dtype=np.int32)
        err = C.create_string_buffer(512)
        self._h = lib().oracle_pptb_build({"ACtree2": 0, "Twobit": 1}[algo], self.M,
                                          _ptr(tb), _ptr(tbw), _ptr(excl),
                                          _ptr(self.high2low), err, 512)
        if not self._h:
            raise PortError(err.value.decode())
        self.max_work = DEFAULT_MAX_WORK

    def blank_dups(self):
        lib().oracle_pptb_blank_dups(self._h)

    def close(self):
        if getattr(self, "_h", None):
            lib().oracle_pptb_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- one subject -------------------------------------------------------------
    def _match(self, head, tail, subject, max_mismatch, min_mismatch, fixed, want_hits):
        hc, hw = (None, None) if head is None else (_u8(head[0]), _i32(head[1]))
        tc, tw = (None, None) if tail is None else (_u8(tail[0]), _i32(tail[1]))
        s = _u8(subject)
        counts = np.zeros(self.M, dtype=np.int32)
        err = C.create_string_buffer(512)
        pk, pe, nh = C.c_void_p(), C.c_void_p(), C.c_long()
        rc = lib().oracle_match(self._h, _ptr(hc), _ptr(hw), _ptr(tc), _ptr(tw),
                                _ptr(s), len(s), max_mismatch, min_mismatch,
                                int(fixed[0]), int(fixed[1]), self.max_work,
                                _ptr(counts),
                                C.byref(pk) if want_hits else None,
                                C.byref(pe) if want_hits else None,
                                C.byref(nh) if want_hits else None, err, 512)
        if rc != 0:
            raise PortError(err.value.decode())
        if not want_hits:
            return counts, None, None
        n = nh.value
        if n:
            keys = np.ctypeslib.as_array(C.cast(pk, C.POINTER(C.c_int)), (n,)).copy()
            ends = np.ctypeslib.as_array(C.cast(pe, C.POINTER(C.c_int)), (n,)).copy()
        else:
            keys = np.zeros(0, np.int32)
            ends = np.zeros(0, np.int32)
        lib().oracle_free(pk)
        lib().oracle_free(pe)
        return counts, keys, ends

    def _shape(self, mode, counts, keys, ends):
        if mode == MATCHES_AS_COUNTS:
            return counts
        if mode == MATCHES_AS_WHICH:       # src/match_reporting.c:150-161
            return (np.nonzero(counts)[0] + 1).astype(np.int32)
        if mode == MATCHES_AS_ENDS:        # src/match_reporting.c:192-201 (NULL if empty)
            out = [None] * self.M
            order = np.argsort(keys, kind="stable")
            ks, es = keys[order], ends[order]
            bounds = np.flatnonzero(np.diff(ks)) + 1
            for seg_k, seg_e in zip(np.split(ks, bounds), np.split(es, bounds)):
                if len(seg_k):
                    out[int(seg_k[0])] = seg_e.astype(np.int32)
            return out
        raise ValueError(mode)

    def match_xstring(self, head, tail, subject, max_mismatch, min_mismatch, fixed, mode,
                      keep=False):
        c, k, e = self._match(head, tail, subject, max_mismatch, min_mismatch, fixed,
                              mode == MATCHES_AS_ENDS)
        return self._shape(mode, c, k, e)

    # ---- views: src/match_pdict.c:225-264 -----------------------------------------
    def match_views(self, head, tail, subject, views_start, views_width, max_mismatch,
                    min_mismatch, fixed, mode, keep=False):
        s = _u8(subject)
        counts = np.zeros(self.M, dtype=np.int64)
        allk, alle = [], []
        for st, wd in zip(np.asarray(views_start).tolist(), np.asarray(views_width).tolist()):
            off = st - 1
            if off < 0 or off + wd > len(s):
                raise PortError("'subject' has \"out of limits\" views")
            c, k, e = self._match(head, tail, s[off:off + wd], max_mismatch, min_mismatch,
                                  fixed, mode == MATCHES_AS_ENDS)
            counts += c
            if k is not None:
                allk.append(k)
                alle.append(e + off)
        keys = np.concatenate(allk) if allk else np.zeros(0, np.int32)
        ends = np.concatenate(alle) if alle else np.zeros(0, np.int32)
        return self._shape(mode, counts.astype(np.int32), keys, ends)

    # ---- sets: src/match_pdict.c:320-346, 386-432, 85-127 ---------------------------
    def vmatch_set(self, head, tail, subj_concat, subj_widths, max_mismatch,
                   min_mismatch, fixed, collapse, weight, mode):
        sc, sw = _u8(subj_concat), _i32(subj_widths)
        offs = np.concatenate([[0], np.cumsum(sw, dtype=np.int64)])
        N = len(sw)
        if mode == MATCHES_AS_WHICH:
            out = []
            for j in range(N):
                c, _, _ = self._match(head, tail, sc[offs[j]:offs[j + 1]], max_mismatch,
                                      min_mismatch, fixed, False)
                out.append((np.nonzero(c)[0] + 1).astype(np.int32))
            return out
        if mode != MATCHES_AS_COUNTS:
            raise PortError("vmatchPDict() is not supported yet, sorry")
        w = np.asarray(weight)
        is_int = w.dtype.kind != "f"
        if collapse == 0:
            ans = np.zeros((self.M, N), dtype=np.int32, order="F")
        elif collapse == 1:
            ans = np.zeros(self.M, dtype=np.int32 if is_int else np.float64)
        elif collapse == 2:
            ans = np.zeros(N, dtype=np.int32 if is_int else np.float64)
        else:
            raise PortError("'collapse' must be FALSE, 1 or 2")
        for j in range(N):
            c, _, _ = self._match(head, tail, sc[offs[j]:offs[j + 1]], max_mismatch,
                                  min_mismatch, fixed, False)
            if collapse == 0:
                ans[:, j] = c
            elif collapse == 1:
                # ans[i] += cnt * weight[j], j outer / i inner (int arithmetic wraps in C)
                if is_int:
                    ans += (c.astype(np.int64) * int(w[j])).astype(np.int32)
                else:
                    ans += c * float(w[j])
            else:
                if is_int:
                    ans[j] = int((c.astype(np.int64) * w.astype(np.int64)).sum())
                else:
                    acc = 0.0
                    for i in np.nonzero(c)[0]:       # zero terms add exactly 0.0
                        acc += int(c[i]) * float(w[i])
                    ans[j] = acc
        return ans
