"""Loaders for the committed golden fixtures (tests/golden/*.npz, written by make_golden.py)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_path(name):
    return name if os.path.isabs(name) else os.path.join(GOLDEN_DIR, name)


def unpack_dynamic(bits, shape, num=None):
    shape = [int(v) for v in shape]
    if num is not None:
        bits = bits[:num]
        shape[0] = bits.shape[0]
    per = shape[1] * shape[2]
    d = np.unpackbits(bits, axis=1)[:, :per]
    return np.ascontiguousarray(d.reshape(shape).astype(np.float32))


def load_inputs(name, num=None):
    """-> (static f32 [B,1+dim,S], dynamic f32 [B,3n,S]) exactly as PACKDataset built them."""
    z = np.load(golden_path(name))
    static = z["static_u8"][:num].astype(np.float32)
    dynamic = unpack_dynamic(z["dynamic_bits"], z["dynamic_shape"], num)
    return np.ascontiguousarray(static), dynamic


def load_traj(name):
    z = np.load(golden_path(name if name.endswith(".npz") else name + ".npz"))
    return {k: z[k] for k in z.files}


def load_rolling(name, num=None):
    """-> dict(adj u8 [B,5,T,T] (adj[b,g,u,v] = edge u -> v), blocks i32 [B,R*T,dim], positions i32 [B,T,dim], T, dim)
    exactly as rolling.RollingDataset / generate.InitialContainer built them."""
    z = np.load(golden_path(name))
    shape = [int(v) for v in z["adj_shape"]]
    bits = z["adj_bits"][:num]
    shape[0] = bits.shape[0]
    per = shape[1] * shape[2] * shape[3]
    adj = np.unpackbits(bits, axis=1)[:, :per].reshape(shape)
    return dict(adj=np.ascontiguousarray(adj), blocks=z["blocks_u8"][:num].astype(np.int32),
                positions=z["positions"][:num].astype(np.int32), T=int(z["total_blocks_num"]), dim=int(z["obj_dim"]))
