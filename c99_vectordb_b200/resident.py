"""Resident index service: a long-lived process keeps the flat indexes in HBM so
that memo's one-shot CLI processes (`main`, memo_cli.py:883-949) stop paying CUDA start-up, the
`.memo` file read and the host->device upload on every invocation (`load_index`, memo_cli.py:251-261).

Two halves, both behind the same faiss-shaped surface memo already uses (memo_cli.py:245-292,
:361, :448):

* the SERVICE (`python -m c99_vectordb_b200.resident serve`) owns the device indexes.  It caches
  one index per `.memo` path, validated by (mtime_ns, size) of the file, so `faiss.read_index(path)`
  from a new CLI process attaches to rows that are already resident; `write_index` re-registers the
  path; an index mutated but never written is dropped on the next open (the file is the truth, as
  in the reference where every process re-reads it).
* the CLIENT classes below (`IndexIDMap2`, `IndexHNSWFlat`, `IndexFlatIP/L2`, `read_index`,
  `write_index`, `vector_to_array`) forward each call over an AF_UNIX stream socket (mode 0600).
  `c99_vectordb_b200/shim_resident/faiss/` makes `import faiss` (memo_cli.py:13) resolve to them.

Wire format: 4-byte magic, u32 header length, JSON header, then the raw bytes of the numpy arrays
the header describes (dtype/shape) — no pickle, nothing executable crosses the socket.

The service holds indexes of `--backend` (default `c99_vectordb_b200.index`, the CUDA index; there
is no CPU fallback — without a device every compute request fails loudly and the error text is
re-raised in the client as RuntimeError, which memo's `load_index` treats like any unreadable file).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import socket
import socketserver
import struct
import sys
import threading
import time
from types import SimpleNamespace

import numpy as np

MAGIC = b"B2RS"
METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1
_DTYPES = {"float32": np.float32, "int64": np.int64, "uint8": np.uint8, "bool": np.bool_}
_MAX_HEADER = 1 << 20


def _private_dir(path: str) -> str:
    """A directory only this user can enter (created 0700; an existing one must be ours and closed to others —
    a predictable name in /tmp could otherwise be squatted by another local user)."""
    try:
        os.mkdir(path, 0o700)
    except FileExistsError:
        pass
    st = os.lstat(path)
    import stat as _stat

    if not _stat.S_ISDIR(st.st_mode) or st.st_uid != os.getuid() or (st.st_mode & 0o077):
        raise PermissionError(f"{path} is not a private directory of uid {os.getuid()}; set B200_RESIDENT_SOCKET")
    return path


def default_socket_path() -> str:
    explicit = os.environ.get("B200_RESIDENT_SOCKET")
    if explicit:
        return explicit
    run = os.environ.get("XDG_RUNTIME_DIR")  # per-user and 0700 by specification
    if run and os.path.isdir(run):
        return os.path.join(run, f"b200-resident-{os.getuid()}.sock")
    return os.path.join(_private_dir(f"/tmp/b200-resident-{os.getuid()}"), "resident.sock")


def _peer_uid(sock: socket.socket) -> int | None:
    """uid of the process on the other end of a unix socket (SO_PEERCRED; None where unsupported)."""
    try:
        creds = sock.getsockopt(socket.SOL_SOCKET, socket.SO_PEERCRED, struct.calcsize("3i"))
        return struct.unpack("3i", creds)[1]
    except (AttributeError, OSError):
        return None


# ------------------------------------------------------------------------------------------------
# framing
# ------------------------------------------------------------------------------------------------
def _recv_exact(sock: socket.socket, n: int, into: memoryview | None = None) -> bytes | None:
    buf = into if into is not None else memoryview(bytearray(n))
    got = 0
    while got < n:
        r = sock.recv_into(buf[got:], n - got)
        if r == 0:
            if got == 0 and into is None:
                return None
            raise ConnectionError("peer closed the connection mid-frame")
        got += r
    return bytes(buf) if into is None else b""


def send_frame(sock: socket.socket, header: dict, arrays=()) -> None:
    arrs = [np.ascontiguousarray(a) for a in arrays]
    for a in arrs:
        if a.dtype.name not in _DTYPES:
            raise TypeError(f"dtype {a.dtype} does not travel")
    header = dict(header, arrays=[{"dtype": a.dtype.name, "shape": list(a.shape)} for a in arrs])
    h = json.dumps(header).encode()
    sock.sendall(MAGIC + struct.pack("<I", len(h)) + h)
    for a in arrs:
        if a.nbytes:
            sock.sendall(memoryview(a).cast("B"))


def recv_frame(sock: socket.socket):
    """-> (header, [arrays]) or None on a clean close between frames."""
    head = _recv_exact(sock, 8)
    if head is None:
        return None
    if head[:4] != MAGIC:
        raise ConnectionError("bad magic")
    (hl,) = struct.unpack("<I", head[4:])
    if hl > _MAX_HEADER:
        raise ConnectionError("header too large")
    header = json.loads(_recv_exact(sock, hl))
    arrays = []
    for spec in header.get("arrays", []):
        dt = _DTYPES[spec["dtype"]]
        shape = tuple(int(s) for s in spec["shape"])
        if any(s < 0 for s in shape):
            raise ConnectionError("negative dimension")
        a = np.empty(shape, dtype=dt)
        if a.nbytes:
            _recv_exact(sock, a.nbytes, memoryview(a).cast("B"))
        arrays.append(a)
    return header, arrays


# ------------------------------------------------------------------------------------------------
# service
# ------------------------------------------------------------------------------------------------
class _Entry:
    __slots__ = ("index", "path", "stamp", "dirty", "last_used", "holders")

    def __init__(self, index, path=None, stamp=None):
        self.index, self.path, self.stamp, self.dirty, self.last_used = index, path, stamp, False, time.monotonic()
        self.holders = 0  # connections that hold a handle to this entry


def _stamp(path: str):
    st = os.stat(path)
    return (st.st_mtime_ns, st.st_size)


class ResidentService:
    """State + request handlers.  One lock: requests are serialised, like calls on one C-ABI handle."""

    def __init__(self, backend="c99_vectordb_b200.index", max_resident: int = 8):
        self.backend = importlib.import_module(backend) if isinstance(backend, str) else backend
        self.max_resident = int(max_resident)
        self.by_path: dict[str, _Entry] = {}
        self.lock = threading.Lock()
        self.stats = {"loads": 0, "hits": 0, "evictions": 0, "requests": 0, "searches": 0, "writes": 0}
        self.last_request = time.monotonic()
        self.stopping = False

    # ---- cache -------------------------------------------------------------------------------
    def _evict(self) -> None:
        while len(self.by_path) > self.max_resident:
            victim = min(self.by_path.values(), key=lambda e: e.last_used)
            del self.by_path[victim.path]
            self.stats["evictions"] += 1

    def _open(self, path: str) -> _Entry:
        stamp = _stamp(path)  # raises FileNotFoundError like read_index on a missing file
        e = self.by_path.get(path)
        if e is not None and not e.dirty and e.stamp == stamp:
            self.stats["hits"] += 1
        else:
            idx = self.backend.read_index(path)
            e = _Entry(idx, path, stamp)
            self.by_path[path] = e
            self.stats["loads"] += 1
            self._evict()
        e.last_used = time.monotonic()
        return e

    # ---- one request -------------------------------------------------------------------------
    def handle(self, table: dict, header: dict, arrays):
        """table: this connection's handle -> _Entry.  Returns (result dict, arrays)."""
        op, a = header["op"], header.get("args", {})
        self.stats["requests"] += 1
        self.last_request = time.monotonic()
        B = self.backend
        if op == "ping":
            return {"pid": os.getpid(), "backend": B.__name__}, ()
        if op == "stats":
            return dict(self.stats, resident=sorted(self.by_path), handles=len(table)), ()
        if op == "shutdown":
            self.stopping = True
            return {}, ()
        if op == "open":
            e = self._open(a["path"])
            return self._register(table, e), ()
        if op == "create":
            base = B.IndexFlat(int(a["d"]), int(a["metric"]))
            idx = B.IndexIDMap2(base) if a.get("idmap", True) else base
            return self._register(table, _Entry(idx)), ()
        e = table.get(int(a["h"]))
        if e is None:
            raise RuntimeError("stale index handle")
        e.last_used = time.monotonic()
        if op in ("add", "reset"):
            e = self._own(table, int(a["h"]), e)  # mutations never show through another connection's handle
        idx = e.index
        if op == "release":
            del table[int(a["h"])]
            e.holders -= 1
            return {}, ()
        if op == "ntotal":
            return {"ntotal": int(idx.ntotal)}, ()
        if op == "ids":
            if not hasattr(idx, "id_map"):
                return {}, (np.arange(int(idx.ntotal), dtype=np.int64),)
            return {}, (np.asarray(B.vector_to_array(idx.id_map), dtype=np.int64),)
        if op == "add":
            x = arrays[0]
            e.dirty = True
            if a.get("with_ids"):
                idx.add_with_ids(x, arrays[1])
            else:
                idx.add(x)
            return {"ntotal": int(idx.ntotal)}, ()
        if op == "reset":
            e.dirty = True
            idx.reset()
            return {}, ()
        if op == "search":
            self.stats["searches"] += 1
            kw = {}
            if a.get("ids_allowed"):
                kw["ids_allowed"] = arrays[1]
            D, I = idx.search(arrays[0], int(a["k"]), **kw)
            return {}, (np.asarray(D, dtype=np.float32), np.asarray(I, dtype=np.int64))
        if op == "write":
            path = a["path"]
            B.write_index(idx, path)
            self.stats["writes"] += 1
            old = self.by_path.get(e.path) if e.path else None
            if old is e and e.path != path:
                del self.by_path[e.path]  # one entry, one file: the index now mirrors `path`
            e.path, e.stamp, e.dirty = path, _stamp(path), False
            self.by_path[path] = e
            self._evict()
            return {}, ()
        raise RuntimeError(f"unknown op {op!r}")

    def _own(self, table: dict, h: int, e: _Entry) -> _Entry:
        """Before a mutation: every process of the reference has a private copy of the index it read, so an unsaved
        add / reset must not be visible to other connections.  A path-cached entry held by this connection alone is
        taken out of the cache and mutated in place (no copy; write() puts it back); one that other connections also
        hold is replaced, for this connection, by a private copy read from the file."""
        if e.path is None or self.by_path.get(e.path) is not e:
            if e.holders <= 1:
                return e  # anonymous or already private
        if e.holders <= 1:
            if e.path is not None and self.by_path.get(e.path) is e:
                del self.by_path[e.path]
            return e
        mine = _Entry(self.backend.read_index(e.path), e.path, _stamp(e.path))
        self.stats["loads"] += 1
        e.holders -= 1
        mine.holders = 1
        table[h] = mine
        return mine

    @staticmethod
    def _register(table: dict, e: _Entry) -> dict:
        h = max(table, default=0) + 1
        table[h] = e
        e.holders += 1
        idx = e.index
        metric = getattr(idx, "metric_type", getattr(getattr(idx, "index", None), "metric_type", METRIC_L2))
        return {"h": h, "d": int(idx.d), "metric": int(metric),
                "ntotal": int(idx.ntotal), "idmap": hasattr(idx, "id_map")}


class _Handler(socketserver.BaseRequestHandler):
    def handle(self):
        svc: ResidentService = self.server.service
        table: dict = {}
        sock = self.request
        uid = _peer_uid(sock)
        if uid is not None and uid not in (os.getuid(), 0):
            return  # the socket is 0600 in a private directory; this is the second lock on the same door
        try:
            while not svc.stopping:
                frame = recv_frame(sock)
                if frame is None:
                    break
                header, arrays = frame
                try:
                    with svc.lock:
                        result, out = svc.handle(table, header, arrays)
                    send_frame(sock, {"ok": True, "result": result}, out)
                except (ConnectionError, BrokenPipeError):
                    raise
                except Exception as ex:  # every failure travels back as text; the client raises
                    send_frame(sock, {"ok": False, "error": f"{type(ex).__name__}: {ex}"})
                if svc.stopping:
                    threading.Thread(target=self.server.shutdown, daemon=True).start()
        except (ConnectionError, BrokenPipeError, OSError):
            pass
        finally:
            for e in table.values():
                e.holders -= 1
            table.clear()  # anonymous indexes die with their connection; path-cached ones stay resident


class _Server(socketserver.ThreadingMixIn, socketserver.UnixStreamServer):
    daemon_threads = True
    allow_reuse_address = True


def make_server(socket_path: str, backend="c99_vectordb_b200.index", max_resident: int = 8) -> _Server:
    """Bind (0600) and return the server; the caller runs serve_forever()."""
    if os.path.exists(socket_path):
        try:  # a live service answers; a stale socket file does not
            ResidentClient(socket_path, autostart=False).call("ping")
            raise RuntimeError(f"a resident service already listens on {socket_path}")
        except (ConnectionError, OSError):
            os.unlink(socket_path)
    old = os.umask(0o177)
    try:
        srv = _Server(socket_path, _Handler)
    finally:
        os.umask(old)
    srv.service = ResidentService(backend, max_resident)
    return srv


def serve(socket_path: str, backend="c99_vectordb_b200.index", max_resident: int = 8, idle_seconds: float = 0.0) -> None:
    srv = make_server(socket_path, backend, max_resident)
    try:  # pay device start-up now, not inside the first client's read_index
        warm = srv.service.backend.IndexFlat(8, METRIC_L2)
        getattr(warm, "close", lambda: None)()
    except Exception as ex:  # no device: keep serving, every compute request will fail loudly
        print(f"[b200 resident] device warm-up failed: {ex}", file=sys.stderr)
    if idle_seconds > 0:
        def reaper():
            while not srv.service.stopping:
                time.sleep(min(1.0, idle_seconds / 4))
                if time.monotonic() - srv.service.last_request > idle_seconds:
                    srv.service.stopping = True
                    srv.shutdown()
        threading.Thread(target=reaper, daemon=True).start()
    try:
        srv.serve_forever(poll_interval=0.05)
    finally:
        srv.server_close()
        try:
            os.unlink(socket_path)
        except OSError:
            pass


# ------------------------------------------------------------------------------------------------
# client
# ------------------------------------------------------------------------------------------------
class ResidentClient:
    """One connection to the service.  `autostart` (default: env B200_RESIDENT_AUTOSTART=1) spawns a
    detached service when nothing listens, then waits for its socket."""

    def __init__(self, socket_path: str | None = None, autostart: bool | None = None, timeout: float = 600.0):
        self.path = socket_path or default_socket_path()
        if autostart is None:
            autostart = os.environ.get("B200_RESIDENT_AUTOSTART", "0") == "1"
        self.sock = self._connect(autostart, timeout)

    def _try(self, timeout):
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.settimeout(timeout)
        try:
            s.connect(self.path)
        except OSError:
            s.close()
            raise
        uid = _peer_uid(s)
        if uid is not None and uid != os.getuid():
            s.close()  # queries, vectors and file paths only go to a service of the same user
            raise PermissionError(f"the service on {self.path} runs as uid {uid}, not {os.getuid()}")
        return s

    def _connect(self, autostart, timeout):
        try:
            return self._try(timeout)
        except PermissionError:
            raise
        except OSError as first:
            if not autostart:
                raise ConnectionError(f"no resident service on {self.path}: {first}") from None
        import subprocess

        idle = os.environ.get("B200_RESIDENT_IDLE_SECONDS", "900")
        subprocess.Popen([sys.executable, "-m", "c99_vectordb_b200.resident", "serve", "--socket", self.path,
                          "--idle-seconds", idle],
                         stdin=subprocess.DEVNULL, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                         start_new_session=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        deadline = time.monotonic() + 120.0
        while time.monotonic() < deadline:
            try:
                return self._try(timeout)
            except OSError:
                time.sleep(0.05)
        raise ConnectionError(f"resident service did not come up on {self.path}")

    def call(self, op: str, args: dict | None = None, arrays=()):
        send_frame(self.sock, {"op": op, "args": args or {}}, arrays)
        frame = recv_frame(self.sock)
        if frame is None:
            raise ConnectionError("resident service closed the connection")
        header, out = frame
        if not header.get("ok"):
            raise RuntimeError(header.get("error", "resident service error"))
        return header.get("result", {}), out

    def close(self) -> None:
        try:
            self.sock.close()
        except OSError:
            pass


_client: ResidentClient | None = None


def client() -> ResidentClient:
    """Process-wide connection used by the faiss-shaped classes below."""
    global _client
    if _client is None:
        _client = ResidentClient()
    return _client


def set_client(c: ResidentClient | None) -> None:
    global _client
    _client = c


class Int64Vector:
    """What `index.id_map` returns; vector_to_array() unwraps it (memo_cli.py:268)."""

    def __init__(self, arr: np.ndarray):
        self._a = arr

    def size(self) -> int:
        return int(self._a.shape[0])

    def at(self, i: int) -> int:
        return int(self._a[i])

    def __len__(self) -> int:
        return self.size()


def vector_to_array(v) -> np.ndarray:
    return np.array(v._a if isinstance(v, Int64Vector) else v, dtype=np.int64, copy=True)


class Index:
    """Client-side proxy of one index held by the service.  Created lazily on first use so that
    memo's `IndexIDMap2(IndexHNSWFlat(d, 32))` (memo_cli.py:245-248) costs one round trip."""

    _idmap = False

    def __init__(self, d: int, metric: int):
        self.d, self.metric_type, self.is_trained = int(d), int(metric), True
        self._h = None
        self._c = None

    def _attach(self, c: ResidentClient, info: dict):
        self._c, self._h = c, int(info["h"])
        self.d, self.metric_type = int(info["d"]), int(info["metric"])
        return self

    def _ensure(self):
        if self._h is None:
            c = client()
            info, _ = c.call("create", {"d": self.d, "metric": self.metric_type, "idmap": self._idmap})
            self._attach(c, info)
        return self._c, self._h

    def __del__(self):
        try:
            if self._h is not None and self._c is not None:
                self._c.call("release", {"h": self._h})
        except Exception:
            pass

    @property
    def ntotal(self) -> int:
        if self._h is None:
            return 0
        return int(self._c.call("ntotal", {"h": self._h})[0]["ntotal"])

    def train(self, x) -> None:
        pass

    def _coerce_x(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d, f"expected [n,{self.d}] got {x.shape}"  # as faiss's wrapper
        return x

    def add(self, x) -> None:
        c, h = self._ensure()
        c.call("add", {"h": h}, (self._coerce_x(x),))

    def search(self, x, k: int, ids_allowed=None):
        c, h = self._ensure()
        k = int(k)
        assert k > 0
        arrays = [self._coerce_x(x)]
        if ids_allowed is not None:
            arrays.append(np.ascontiguousarray(ids_allowed, dtype=np.int64).reshape(-1) if isinstance(ids_allowed, np.ndarray)
                          else np.fromiter((int(i) for i in ids_allowed), dtype=np.int64))
        _, (D, I) = c.call("search", {"h": h, "k": k, "ids_allowed": ids_allowed is not None}, arrays)
        return D, I

    def reset(self) -> None:
        c, h = self._ensure()
        c.call("reset", {"h": h})


class IndexFlat(Index):
    def __init__(self, d: int, metric: int = METRIC_L2):
        super().__init__(d, metric)


class IndexFlatIP(IndexFlat):
    def __init__(self, d: int):
        super().__init__(d, METRIC_INNER_PRODUCT)


class IndexFlatL2(IndexFlat):
    def __init__(self, d: int):
        super().__init__(d, METRIC_L2)


class IndexHNSWFlat(IndexFlat):
    """Exact flat L2 index with the attribute bag memo writes to (memo_cli.py:246-247)."""

    def __init__(self, d: int, M: int = 32, metric: int = METRIC_L2):
        super().__init__(d, metric)
        self.hnsw = SimpleNamespace(efConstruction=40, efSearch=16, M=int(M))


class IndexIDMap(Index):
    _idmap = True

    def __init__(self, index: Index):
        if not isinstance(index, IndexFlat):
            raise RuntimeError("IndexIDMap: only flat base indexes are supported")
        if index._h is not None:
            raise RuntimeError("index must be empty on input")
        super().__init__(index.d, index.metric_type)
        self.index = index
        self.own_fields = True

    @property
    def id_map(self) -> Int64Vector:
        if self._h is None:
            return Int64Vector(np.zeros((0,), dtype=np.int64))
        return Int64Vector(self._c.call("ids", {"h": self._h})[1][0])

    def add_with_ids(self, x, ids) -> None:
        c, h = self._ensure()
        x = self._coerce_x(x)
        ids = np.ascontiguousarray(ids, dtype=np.int64).reshape(-1)
        assert ids.shape[0] == x.shape[0]
        c.call("add", {"h": h, "with_ids": True}, (x, ids))

    def add(self, x) -> None:
        raise RuntimeError("add does not make sense with IndexIDMap, use add_with_ids")  # as faiss


class IndexIDMap2(IndexIDMap):
    pass


def read_index(path: str) -> Index:
    """faiss.read_index (memo_cli.py:255): attaches to the resident copy when the file is unchanged,
    otherwise the service loads it.  Raises RuntimeError on a missing/corrupt file."""
    c = client()
    info, _ = c.call("open", {"path": os.path.abspath(str(path))})
    cls = IndexIDMap2 if info.get("idmap") else IndexFlat
    idx = cls.__new__(cls)
    Index.__init__(idx, info["d"], info["metric"])
    return idx._attach(c, info)


def write_index(index: Index, path: str) -> None:
    """faiss.write_index (memo_cli.py:361, :448): the service writes the faiss-layout file and keeps
    the index resident under that path."""
    c, h = index._ensure()
    c.call("write", {"h": h, "path": os.path.abspath(str(path))})


# ------------------------------------------------------------------------------------------------
def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m c99_vectordb_b200.resident")
    sub = ap.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("serve")
    s.add_argument("--socket", default=default_socket_path())
    s.add_argument("--backend", default="c99_vectordb_b200.index")
    s.add_argument("--max-resident", type=int, default=8)
    s.add_argument("--idle-seconds", type=float, default=0.0, help="exit after this long without a request (0 = never)")
    for name in ("stats", "stop", "ping"):
        p = sub.add_parser(name)
        p.add_argument("--socket", default=default_socket_path())
    a = ap.parse_args(argv)
    if a.cmd == "serve":
        serve(a.socket, a.backend, a.max_resident, a.idle_seconds)
        return 0
    c = ResidentClient(a.socket, autostart=False)
    print(json.dumps(c.call({"stop": "shutdown"}.get(a.cmd, a.cmd))[0]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
