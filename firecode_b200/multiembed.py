"""Batched replacement of ``firecode.multiembed`` (multiembed.py:23-159): every arrangement of interacting atom
pairs of a bifunctional + bifunctional system is embedded IN THIS PROCESS, one after the other on one CUDA
context, and the results are concatenated in ARRANGEMENT ORDER.

The reference starts one child ``Embedder`` per arrangement in a ``ProcessPoolExecutor`` (multiembed.py:58-70) --
with a GPU library that would mean one CUDA context per child -- and collects them with ``as_completed``, so its
output order depends on scheduling.  Here a child is still what the reference builds (a cyclical embed of the two
molecules with the pairings ``x`` / ``y`` on the chosen atoms, then compenetration, fitness and similarity refining
with ``rmsd=False``, multiembed.py:145-149), but its screens are this package's: ``embeds.cyclical_embed`` and the
``refining`` steps.  Children come from ``make_child`` -- by default the host application's own ``Embedder`` on a
generated input file, exactly as ``run_child_embedder`` writes it (multiembed.py:102-143); tests pass a factory
that rebuilds the children from stored arrays.
"""

from __future__ import annotations

import os
import time
from itertools import permutations

import numpy as np

from . import embeds, refining
from .errors import ZeroCandidatesError
from .utils import cartesian_product


class InputError(Exception):
    """firecode.errors.InputError stand-in when FIRECODE itself is not importable."""


def arrangements(reactive1, reactive2):
    """Every arrangement of two interacting pairs not insisting on the same atom twice, in the reference's order
    (multiembed.py:39-48): ``pairs = cartesian_product(...)`` then ``permutations(pairs, 2)`` filtered."""
    pairs = cartesian_product(np.asarray(reactive1), np.asarray(reactive2))
    return [((int(ix_1), int(ix_2)), (int(iy_1), int(iy_2)))
            for ((ix_1, ix_2), (iy_1, iy_2)) in permutations(pairs, 2) if ix_1 != iy_1 and ix_2 != iy_2]


def host_child(embedder, arrangement, i):
    """The child embedder of arrangement ``i`` built by the host application (FIRECODE must be importable): same
    folder, files and input line as run_child_embedder (multiembed.py:102-143).  Returns (child, cleanup)."""
    from shutil import copy, rmtree

    from firecode.embedder import Embedder, RunEmbedding

    mol1, mol2 = embedder.objects
    options = embedder.options
    (ix_1, ix_2), (iy_1, iy_2) = arrangement
    parent = os.getcwd()
    folder = os.path.join(parent, f"firecode_embed{i + 1}")
    os.makedirs(folder, exist_ok=True)
    os.chdir(folder)
    try:
        copy(os.path.join(parent, mol1.filename), mol1.filename)
        copy(os.path.join(parent, mol2.filename), mol2.filename)
        child_name = f"embed{i + 1}_input.txt"
        with open(child_name, "w", encoding="utf-8") as f:
            extra = ""
            extra += " debug" if options.debug else ""
            extra += " simpleorbitals" if options.simpleorbitals else ""
            extra += f" shrink={options.shrink_multiplier}" if options.shrink else ""
            f.write(f"noopt {extra}\n")
            f.write(f"{mol1.filename} {ix_1}x {iy_1}y\n")
            f.write(f"{mol2.filename} {ix_2}x {iy_2}y\n")
        child = RunEmbedding(Embedder(os.path.join(os.getcwd(), child_name), stamp=f"embed{i + 1}"))
        child._set_reactive_atoms_cumnums()
        child.write_mol_info()
    except BaseException:
        os.chdir(parent)
        raise

    def cleanup():
        os.chdir(parent)
        if not options.debug:
            rmtree(folder, ignore_errors=True)

    return child, cleanup


def run_child(child):
    """What run_child_embedder does with a child (multiembed.py:135-153), through this package's screens.
    Returns (structures, constrained_indices); structures is empty when the child has no candidates."""
    try:
        child.structures = embeds.cyclical_embed(child)
        if hasattr(child, "objects") and all(hasattr(m, "atoms") for m in child.objects):
            child.atoms = np.concatenate([m.atoms for m in child.objects])
        refining.compenetration_refining(child)
        refining.fitness_refining(child)
        refining.similarity_refining(child, rmsd=False, verbose=True)
    except ZeroCandidatesError:
        child.structures = np.array([], dtype=float)
    return child.structures, getattr(child, "constrained_indices", np.zeros((0, 2, 2), dtype=np.int64))


def multiembed_bifunctional(embedder, make_child=None):
    """Drop-in for firecode.multiembed.multiembed_bifunctional (multiembed.py:33-100).  Sets embedder.structures,
    embedder.atoms, embedder.constrained_indices and returns the structures, children in arrangement order."""
    mol1, mol2 = embedder.objects
    todo = arrangements(mol1.reactive_indices, mol2.reactive_indices)
    make_child = host_child if make_child is None else make_child
    structures_out, constr_ids = [], []
    embedder.t_start_run = time.perf_counter()
    embedder.log()
    embedder.log(f"--> Multiembed: running {len(todo)} embeds in one process (arrangement order)")
    per_child = []
    for i, arrangement in enumerate(todo):
        t0 = time.perf_counter()
        made = make_child(embedder, arrangement, i)
        child, cleanup = made if isinstance(made, tuple) else (made, None)
        try:
            structures, constrained = run_child(child)
        finally:
            if cleanup is not None:
                cleanup()
        embedder.log(f"--> Child embed {i + 1:3}/{len(todo):3}: generated {len(structures):4} candidates in "
                     f"{time.perf_counter() - t0:.3f} s.")
        per_child.append(len(structures))
        if len(structures) > 0:
            structures_out.append(np.asarray(structures))
            constr_ids.append(np.asarray(constrained))
    embedder.b200_multiembed_counts = per_child
    if not structures_out:
        raise ZeroCandidatesError("--> Multiembed did not find any suitable disposition of molecules.")
    embedder.structures = np.concatenate(structures_out)
    embedder.atoms = np.concatenate([mol.atoms for mol in embedder.objects])
    # only the interaction constraints: the internal ones are added later, during refinement (multiembed.py:86-87)
    embedder.constrained_indices = np.concatenate(constr_ids)
    if hasattr(embedder, "write_structures"):
        embedder.write_structures("embedded", energies=False)
    embedder.log(f"\n--> Multiembed completed: generated {len(embedder.structures)} candidates from "
                 f"{len(structures_out)} arrangements in {time.perf_counter() - embedder.t_start_run:.3f} s.")
    return embedder.structures


def multiembed_dispatcher(embedder):
    """firecode.multiembed.multiembed_dispatcher (multiembed.py:23-30)."""
    if len(embedder.objects) == 2:
        return multiembed_bifunctional(embedder)
    try:
        from firecode.errors import InputError as HostInputError
    except Exception:  # FIRECODE not importable: same message, local exception type
        HostInputError = InputError
    raise HostInputError("The multiembed requested is currently unavailable.")
