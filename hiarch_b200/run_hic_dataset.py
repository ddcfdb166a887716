"""Mirror of the reference's batch harness (publish/run_hic_dataset.py): map a function over a
list of sample files, sharded over worker processes -- here one worker per GPU.

    ParallelFunc(species, dataset_name, func, mtx_files={name: (mtx, bed)}).vali_parallel(out, thread)

`vali_parallel` splits the file list into contiguous shards with the reference's own
`yield_mp_idxes` (:259-269), starts one process per shard (spawn, so CUDA is initialised after
the process starts -- the reference forks, :587), pins shard i to GPU i mod n_gpus, runs
`func(mtx, **kwargs)` on every file, harvests '#para-name: value' lines from stdout like
CaptureParas (:314-356) and concatenates the per-file result tables into one TSV (:590-598).
No communication between workers.
"""
import io
import os
import sys

import pandas as pd

from .dist import yield_mp_idxes  # noqa: F401  (public name of the reference)


def read_mtx(mtx_file, index_file=None, **kwargs):
    """read_mtx (:19-93) for the '.mtx' text format (the .mcool / bedtools paths need external
    tools that are absent: out of scope)."""
    from .new_hic_class import read_matrix
    return read_matrix(mtx_file, index_file)


class CaptureParas(io.StringIO):
    """stdout tee that harvests lines starting with '#para-' (:314-356)."""

    def __init__(self):
        super().__init__()
        self.paras = {}

    def write(self, s):
        for line in s.splitlines():
            if line.startswith('#para-') and ':' in line:
                k, v = line[len('#para-'):].split(':', 1)
                self.paras[k.strip()] = v.strip()
        return len(s)


class PerCell:
    """Picklable per-file keyword value: `template.format(cell=name)` (what update_para_output
    stores; a lambda would not survive the pickling of the spawned workers' arguments)."""

    def __init__(self, template):
        self.template = template

    def __call__(self, cell):
        return self.template.format(cell=cell)

    def __repr__(self):
        return f"PerCell({self.template!r})"


def _worker(shard_id, file_names, mtx_files, func, func_kwargs, n_gpus, capture_paras, sink):
    import torch
    if n_gpus > 0:
        torch.cuda.set_device(shard_id % n_gpus)
    for name in file_names:
        mtx_file, index_file = mtx_files[name]
        mtx = read_mtx(mtx_file, index_file)
        kwargs = {k: (v(name) if callable(v) else v) for k, v in func_kwargs.items()}
        cap = CaptureParas() if capture_paras else None
        old = sys.stdout
        try:
            if cap is not None:
                sys.stdout = cap
            result = func(mtx.copy(), **kwargs)
        finally:
            sys.stdout = old
        if isinstance(result, pd.DataFrame):
            df = result.copy()
        elif isinstance(result, dict):
            df = pd.DataFrame([result])
        else:
            df = pd.DataFrame([{'value': result}])
        df.insert(0, 'cell', name)
        if cap is not None:
            for k, v in cap.paras.items():
                df[k] = v
        sink.append(df)


class ParallelFunc:
    def __init__(self, species, dataset_name, func, spread_type=None, read_type='den', n_cell=None,
                 include_pats=None, exclude_pats=None, spread_include_pat=None,
                 out_concat_pre_value=False, mtx_files=None):
        if mtx_files is None:
            raise ValueError("pass mtx_files={name: (mtx_file, index_file)}: the reference's "
                             "directory-layout globbing (data_dir/species/mtx/dataset) is out of scope")
        self.species, self.dataset_name = species, dataset_name
        self.func = func
        self.func_kwargs = {}
        self.file_names = sorted(mtx_files)
        if include_pats is not None:
            pats = [include_pats] if isinstance(include_pats, str) else include_pats
            self.file_names = [f for f in self.file_names if any(p in f for p in pats)]
        if exclude_pats is not None:
            pats = [exclude_pats] if isinstance(exclude_pats, str) else exclude_pats
            self.file_names = [f for f in self.file_names if not any(p in f for p in pats)]
        if n_cell is not None:
            self.file_names = self.file_names[:n_cell]
        self.mtx_files = {f: mtx_files[f] for f in self.file_names}

    def update_func_kwargs(self, new_dict):
        """Values may be callables of the cell name, like the reference (:403-406)."""
        self.func_kwargs.update(new_dict)

    def update_para_output(self, output, output_para='output'):
        self.func_kwargs[output_para] = PerCell(str(output).replace('{', '{{').replace('}', '}}') + '/{cell}')

    def shards(self, thread):
        return [[self.file_names[i] for i in idx] for idx in yield_mp_idxes(len(self.file_names), thread)]

    def vali_parallel(self, out_file=None, thread=2, capture_paras=True):
        """:573-603. `thread` = worker processes (reference default 2); None or 0 = one per visible GPU.
        Worker i runs on GPU i mod n_gpus."""
        import torch
        import torch.multiprocessing as mp
        n_gpus = torch.cuda.device_count() if torch.cuda.is_available() else 0
        thread = thread or max(1, n_gpus)
        shards = self.shards(thread)
        if len(shards) == 1 or thread == 1:
            sink = []
            for i, names in enumerate(shards):
                _worker(i, names, self.mtx_files, self.func, self.func_kwargs, n_gpus, capture_paras, sink)
            results = sink
        else:
            import pickle
            for name, v in self.func_kwargs.items():      # fail early and clearly, not inside a worker
                try:
                    pickle.dumps(v)
                except Exception as e:
                    raise TypeError(f"func_kwargs[{name!r}] cannot be sent to the worker processes ({e}); use a "
                                    "module-level function or run_hic_dataset.PerCell('...{cell}...')") from e
            ctx = mp.get_context("spawn")
            with ctx.Manager() as manager:
                sink = manager.list()
                procs = [ctx.Process(target=_worker, args=(i, names, self.mtx_files, self.func,
                                                           self.func_kwargs, n_gpus, capture_paras, sink))
                         for i, names in enumerate(shards)]
                for p in procs:
                    p.start()
                for p in procs:
                    p.join()
                bad = [p.exitcode for p in procs if p.exitcode != 0]
                if bad:
                    raise RuntimeError(f"{len(bad)} worker(s) failed: exit codes {bad}")
                results = list(sink)
        table = pd.concat(results, axis=0, ignore_index=True) if results else pd.DataFrame([])
        if len(table):
            table = table.sort_values('cell', kind='stable').reset_index(drop=True)
        if out_file is not None:
            table.to_csv(out_file, sep="\t", index=False)
        return table
