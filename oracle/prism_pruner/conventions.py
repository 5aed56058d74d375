"""Switches for the prism_pruner conventions that /root/reference cannot pin.

The CUDA product exposes the same switches (firecode_b200.conventions) and the parity tests run
both sides with identical settings.
"""

# rot_mat_from_pointer: +1 = right-handed (counter-clockwise about the pointer, scipy
# Rotation.from_rotvec convention), -1 = left-handed.
ROT_HANDEDNESS = +1

# rotate_dihedral: the rotation axis is coords[i2]-coords[i3] (+1) or coords[i3]-coords[i2] (-1).
TORSION_AXIS_SIGN = +1

# prune(): which member of a similar pair is dropped, and whether a pass reads a snapshot of the
# mask taken at its start ("snapshot") or the mask as it is being updated ("greedy").
# BASELINE.json north_star specifies keep-first; upstream recollection is ("last", "snapshot").
PRUNE_KEEP = "first"
PRUNE_PASS_MODE = "greedy"

# prune(): minimum average number of active structures per chunk for a k-pass to run.
PRUNE_MIN_PER_CHUNK = 20
PRUNE_CHUNK_OVER = "full"   # a pass cuts the "full" array in k chunks of n // k structures, or the "active" structures in k chunks of n_active // k

# prune_by_rmsd(): heavy atoms only, centred Kabsch, max deviation default = 2 * max_rmsd
PRUNE_RMSD_HEAVY_ONLY = True
PRUNE_MAXDEV_FACTOR = 2.0

# prune_by_moment_of_inertia(): relative deviation allowed on each principal moment
MOI_MAX_DEVIATION = 1e-2
