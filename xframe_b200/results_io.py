"""Result wire format of `fxs reconstruct`.

Inside xFrame the record of `post_processing` (reconstruct.py:160-183, built here by worker.assemble_reconstruction_record) is handed to the
reference's own `db.save('reconstructions', record)`, which writes `data.h5` with its generic nested-dict rules
(externalLibraries/hdf5_plugin.py:53-131): dict -> group, list / tuple -> group with attrs['type'] = 'list' / 'tuple' and children '0', '1', ...,
ndarray -> dataset (complex as '<c16'), scalar -> dataset, str -> utf-8 bytes dataset with attrs['type'] = 'str'.

Outside xFrame this module writes the SAME tree: to `data.h5` with those rules when h5py is importable, otherwise to `data.npz` (one array per
leaf, keyed by its HDF5 path, plus a `__tree__` JSON entry holding the list / tuple / str / scalar markers) -- h5py is not part of this
image.  `load_reconstructions` restores the nested record from either file, so downstream consumers (`fxs average`) see the layout of
projects/fxs/_database_.py:223-390 and tests/test_fxs_integration.py:388-421.
"""
import json
import os

import numpy as np


def _walk(node, path, leaves, marks):
    if isinstance(node, dict):
        if not node and path:
            marks[path] = 'dict:empty'
        for k, v in node.items():
            _walk(v, f'{path}/{k}' if path else str(k), leaves, marks)
    elif isinstance(node, (list, tuple)):
        marks[path] = 'list' if isinstance(node, list) else 'tuple'
        for i, v in enumerate(node):
            _walk(v, f'{path}/{i}', leaves, marks)
        if len(node) == 0:
            marks[path] += ':empty'
    elif isinstance(node, str):
        marks[path] = 'str'
        leaves[path] = np.frombuffer(node.encode('utf-8'), dtype=np.uint8).copy()
    elif node is None:
        marks[path] = 'none'
        leaves[path] = np.zeros(0)
    elif isinstance(node, np.ndarray):
        leaves[path] = node.astype('<c16') if np.iscomplexobj(node) else node
    elif isinstance(node, (complex, float, int, bool, np.number, np.bool_)):
        marks[path] = 'scalar'
        leaves[path] = np.asarray(node)
    else:
        raise ValueError(f'cannot save {type(node)} at {path}')


def save_reconstructions(record, run_path):
    """Write `record` under directory run_path; returns the file written."""
    os.makedirs(run_path, exist_ok=True)
    leaves, marks = {}, {}
    _walk(record, '', leaves, marks)
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:
        out = os.path.join(run_path, 'data.h5')
        with h5py.File(out, 'w') as f:
            for path, mark in marks.items():
                if mark.startswith(('list', 'tuple')):
                    f.require_group(path).attrs['type'] = mark.split(':')[0]
                elif mark == 'dict:empty':
                    f.require_group(path)
            for path, arr in leaves.items():
                mark = marks.get(path)
                if mark == 'str':
                    ds = f.create_dataset(path, data=arr.tobytes())
                    ds.attrs['type'] = 'str'
                elif mark == 'scalar':
                    f.create_dataset(path, data=arr[()])
                else:
                    f.create_dataset(path, data=arr)
        return out
    out = os.path.join(run_path, 'data.npz')
    np.savez_compressed(out, __tree__=np.frombuffer(json.dumps(marks).encode(), dtype=np.uint8), **{p.replace('/', '|'): a for p, a in leaves.items()})
    return out


def _insert(tree, path, value):
    keys = path.split('/')
    for k in keys[:-1]:
        tree = tree.setdefault(k, {})
    tree[keys[-1]] = value


def _finish(node, path, marks):
    if not isinstance(node, dict):
        return node
    out = {k: _finish(v, f'{path}/{k}' if path else k, marks) for k, v in node.items()}
    mark = marks.get(path, '')
    if mark.startswith(('list', 'tuple')):
        seq = [out[str(i)] for i in range(len(out))]
        return seq if mark.startswith('list') else tuple(seq)
    return out


def load_reconstructions(file_path):
    """Nested record from `data.npz` (this module) or `data.h5` (this module or the reference's db.save; needs h5py)."""
    if file_path.endswith('.npz'):
        z = np.load(file_path, allow_pickle=False)
        marks = json.loads(z['__tree__'].tobytes().decode())
        tree = {}
        for key in z.files:
            if key == '__tree__':
                continue
            path = key.replace('|', '/')
            a = z[key]
            mark = marks.get(path)
            if mark == 'str':
                a = a.tobytes().decode('utf-8')
            elif mark == 'scalar':
                a = a[()]
            elif mark == 'none':
                a = None
            _insert(tree, path, a)
        for path, mark in marks.items():
            if mark.endswith(':empty'):
                _insert(tree, path, {})
        return _finish(tree, '', marks)
    import h5py

    def rec(g):
        out = {}
        for k, item in g.items():
            t = item.attrs.get('type', False)
            if isinstance(item, h5py.Dataset):
                out[k] = item[()].decode('utf-8') if t == 'str' else item[()]
            else:
                d = rec(item)
                out[k] = [d[str(i)] for i in range(len(d))] if t == 'list' else tuple(d[str(i)] for i in range(len(d))) if t == 'tuple' else d
        return out
    with h5py.File(file_path, 'r') as f:
        return rec(f)
