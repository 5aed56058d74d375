"""Conventions of the absent ``prism_pruner`` dependency that the reference tree cannot pin
(SURVEY.md 8c).  The CPU oracle (oracle/prism_pruner/conventions.py) holds the same switches and
the parity tests run both sides with identical settings; results are reported with the values used.
"""

ROT_HANDEDNESS = +1        # rot_mat_from_pointer: +1 right-handed about the pointer
TORSION_AXIS_SIGN = +1     # rotate_dihedral axis = sign * (coords[i2] - coords[i3])
PRUNE_KEEP = "first"       # which member of a similar pair survives
PRUNE_PASS_MODE = "greedy"  # "greedy" (mask updated in place) or "snapshot" (mask read at pass start)
PRUNE_MIN_PER_CHUNK = 20
PRUNE_CHUNK_OVER = "full"   # a pass cuts the "full" array in k chunks of n // k structures, or the "active" structures in k chunks of n_active // k
PRUNE_RMSD_HEAVY_ONLY = True
PRUNE_MAXDEV_FACTOR = 2.0
MOI_MAX_DEVIATION = 1e-2


def as_dict():
    return {k: v for k, v in globals().items() if k.isupper()}
