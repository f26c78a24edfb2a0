"""Small helpers of the problem-definition API (reference: fedm/utils.py:6-19)."""
import os


def _rank():
    return int(os.environ.get("RANK", "0"))


def print_rank_0(*args, **kwargs):
    """print() on the first process only (one process per GPU replaces the reference's MPI ranks)."""
    if _rank() == 0 and not os.environ.get("FEDM_B200_QUIET"):
        print(*args, **kwargs)


def comma_separated(items):
    """'a', 'b', 'c' -- used in the API's error messages."""
    return ", ".join(f"'{item}'" for item in items)


def mesh_info(mesh):
    """Element count and edge-length extremes (summed / reduced over processes in the reference)."""
    return (f"Number of elements is: {int(mesh.num_cells())}\n"
            f"Maximum element edge length is: {mesh.hmax():.5g}\n"
            f"Minimum element edge length is: {mesh.hmin():.5g}")
