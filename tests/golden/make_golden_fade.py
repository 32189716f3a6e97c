"""Golden vectors for the real-image preparation, produced by EXECUTING the reference's own
function bodies: dataset.py cannot be imported here (h5py, librosa are absent), so the source of
``DepthDataset.__getitem__``'s helpers -- ``OldH5Dataset.alpha_fade`` (dataset.py:109-113) and
``utils.adjust_dynamic_range`` (utils.py:24-30) -- is cut out of the reference files with ``ast`` and compiled as is.

    python tests/golden/make_golden_fade.py     ->  tests/golden/fade.npz
"""
import ast
import os
import types

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


def function_from(path, name, cls=None):
    src = open(path).read()
    tree = ast.parse(src)
    nodes = tree.body
    if cls is not None:
        nodes = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls][0].body
    fn = [n for n in nodes if isinstance(n, ast.FunctionDef) and n.name == name][0]
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {'np': np}
    exec(compile(mod, path, 'exec'), ns)
    return ns[name]


def main():
    alpha_fade = function_from(os.path.join(REF, 'dataset.py'), 'alpha_fade', cls='OldH5Dataset')
    adjust = function_from(os.path.join(REF, 'utils.py'), 'adjust_dynamic_range')
    rng = np.random.RandomState(7)
    out = {}
    cases = []
    for i, (c, h, w, dtype, alpha, rin, rout) in enumerate([
            (3, 8, 8, 'uint8', 0.25, (0, 255), (-1, 1)), (3, 16, 16, 'uint8', 0.0, (0, 255), (-1, 1)),
            (1, 32, 32, 'uint8', 0.99984, (0, 255), (-1, 1)), (3, 8, 8, 'uint8', 1.0, (0, 255), (-1, 1)),
            (1, 16, 16, 'float32', 0.5, (-1, 1), (-1, 1)), (3, 4, 4, 'float32', 0.3, (0, 1), (-1, 1))]):
        n = 3
        if dtype == 'uint8':
            batch = rng.randint(0, 256, size=(n, c, h, w)).astype(np.uint8)
        else:
            batch = rng.uniform(rin[0], rin[1], size=(n, c, h, w)).astype(np.float32)
        res = []
        for d in batch:
            holder = types.SimpleNamespace(alpha=alpha)
            x = d
            if alpha < 1.0:                       # dataset.py:62-63
                x = alpha_fade(holder, x)
            x = adjust(x, rin, rout)              # dataset.py:65
            res.append(x.astype('float32'))       # dataset.py:67
        out['in%d' % i] = batch
        out['out%d' % i] = np.stack(res)
        cases.append((alpha, rin[0], rin[1], rout[0], rout[1]))
    out['cases'] = np.array(cases, dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, 'fade.npz'), **out)
    print('wrote fade.npz with', len(cases), 'cases')


if __name__ == '__main__':
    main()
