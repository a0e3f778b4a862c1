"""Writers/readers for the file formats on BPtrain's surface, used to synthesise inputs:
QuickNet Pfile (as consumed by reference Interface.cc:468-555, 689-861), norm file (:301-326), MAT-v4 .wts
(:351-391, 411-465; same format Gen_rand_net.cpp:145-170 produces)."""
import struct

import numpy as np

PFILE_HEADER_SIZE = 32768


def write_pfile(path, sentences):
    """sentences: list of float32 arrays [T_i x dim].  Records are (sent_id, frame_id, dim floats), big-endian;
    the tail is the (num_sentences+1) cumulative start-frame index."""
    dim = sentences[0].shape[1]
    n_frames = int(sum(s.shape[0] for s in sentences))
    hdr = (f"-pfile_header version 0 size {PFILE_HEADER_SIZE}\n-num_sentences {len(sentences)}\n"
           f"-num_frames {n_frames}\n-first_feature_column 2\n-num_features {dim}\n-first_label_column {2 + dim}\n"
           f"-num_labels 0\n-format dd{'f' * dim}\n-data size {n_frames * (2 + dim)} offset 0 ndim 2 nrow {n_frames} "
           f"ncol {2 + dim}\n-sent_table_data size {len(sentences) + 1} offset {n_frames * (2 + dim)} ndim 1\n-end\n")
    with open(path, "wb") as f:
        f.write(hdr.encode().ljust(PFILE_HEADER_SIZE, b"\0"))
        for si, s in enumerate(sentences):
            T = s.shape[0]
            rec = np.empty((T, 2 + dim), dtype=">u4")
            rec[:, 0] = si
            rec[:, 1] = np.arange(T)
            rec[:, 2:] = np.ascontiguousarray(s, dtype=">f4").view(">u4")
            f.write(rec.tobytes())
        starts = np.concatenate([[0], np.cumsum([s.shape[0] for s in sentences])]).astype(">u4")
        f.write(starts.tobytes())


def read_pfile(path):
    """Returns (list of float32 arrays [T_i x dim], sentence ids [n_frames], frame ids [n_frames])."""
    raw = open(path, "rb").read()
    hdr = raw[:PFILE_HEADER_SIZE].split(b"\0", 1)[0].decode()
    fields = {ln.split()[0]: ln.split()[1:] for ln in hdr.splitlines() if ln.startswith("-") and len(ln.split()) > 1}
    n_sent, n_frames, dim = int(fields["-num_sentences"][0]), int(fields["-num_frames"][0]), int(fields["-num_features"][0])
    rec = np.frombuffer(raw, dtype=">u4", count=n_frames * (2 + dim), offset=PFILE_HEADER_SIZE).reshape(n_frames, 2 + dim)
    starts = np.frombuffer(raw, dtype=">u4", count=n_sent + 1, offset=PFILE_HEADER_SIZE + 4 * n_frames * (2 + dim))
    assert len(raw) == PFILE_HEADER_SIZE + 4 * (n_frames * (2 + dim) + n_sent + 1)
    data = rec[:, 2:].astype(">u4").view(">f4").astype(np.float32)
    sents = [data[int(starts[i]):int(starts[i + 1])] for i in range(n_sent)]
    return sents, rec[:, 0].astype(np.int64), rec[:, 1].astype(np.int64)


def write_norm(path, mean, inv_std):
    with open(path, "w") as f:
        f.write(f"vec {len(mean)}\n")
        for v in mean:
            f.write(f"{float(v):.9g}\n")
        f.write(f"vec {len(inv_std)}\n")
        for v in inv_std:
            f.write(f"{float(v):.9g}\n")


def write_wts(path, weights, bias):
    """weights[i]: [n_in x n_out] float32 (w[in*n_out+out]) for i = 1..L; MATLAB sees an n_out x n_in single matrix."""
    with open(path, "wb") as f:
        for i in range(1, len(weights)):
            w = np.ascontiguousarray(weights[i], dtype="<f4")
            b = np.ascontiguousarray(bias[i], dtype="<f4")
            n_in, n_out = w.shape
            name = f"weights{i}{i + 1}".encode() + b"\0"
            f.write(struct.pack("<5i", 10, n_out, n_in, 0, len(name)) + name + w.tobytes())
            name = f"bias{i + 1}".encode() + b"\0"
            f.write(struct.pack("<5i", 10, 1, n_out, 0, len(name)) + name + b.tobytes())


def read_wts(path, layersizes):
    ws, bs = [None], [None]
    with open(path, "rb") as f:
        for i in range(1, len(layersizes)):
            t, n_out, n_in, _, nl = struct.unpack("<5i", f.read(20))
            f.read(nl)
            assert (t, n_out, n_in) == (10, layersizes[i], layersizes[i - 1]), (t, n_out, n_in)
            ws.append(np.frombuffer(f.read(4 * n_in * n_out), dtype="<f4").reshape(n_in, n_out).copy())
            t, one, n, _, nl = struct.unpack("<5i", f.read(20))
            f.read(nl)
            assert (t, one, n) == (10, 1, layersizes[i])
            bs.append(np.frombuffer(f.read(4 * n), dtype="<f4").copy())
    return ws, bs


def synth_corpus(n_sent, dim, out_dim, seed=1, min_len=50, max_len=400, ctx_center=None):
    """SURVEY.md §8d synthetic corpus: raw features x ~ N(mu_j, 1.5^2), mu_j = -6 + 3 sin(j/40); norm = exact mu,
    1/sigma; targets y = tanh(A x_hat) + 0.1 eps (already normalised, the reader does not normalise targets)."""
    rng = np.random.default_rng(seed)
    mu = (-6 + 3 * np.sin(np.arange(dim) / 40.0)).astype(np.float32)
    sigma = np.full(dim, 1.5, dtype=np.float32)
    A = (np.random.default_rng(seed + 1).standard_normal((dim, out_dim)) / np.sqrt(dim)).astype(np.float32)
    feas, targs = [], []
    for _ in range(n_sent):
        T = int(rng.integers(min_len, max_len + 1))
        xh = rng.standard_normal((T, dim), dtype=np.float32)
        feas.append((xh * sigma + mu).astype(np.float32))
        targs.append((np.tanh(xh @ A) + 0.1 * rng.standard_normal((T, out_dim), dtype=np.float32)).astype(np.float32))
    return feas, targs, mu, (1.0 / sigma).astype(np.float32)
