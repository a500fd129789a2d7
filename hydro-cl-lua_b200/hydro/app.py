"""Pieces of hydro/app.lua the hot path needs: the limiter table and the ``real`` selection.

Reference: hydro/app.lua:614-635 (limiters, 1-based there, 0-based here), :892,926-929 (real).
"""

# order == hydro/app.lua:615-634; index 0 ('donor cell') means "no flux limiter" (fvsolver.lua:61-63)
limiterNames = [
    "donor cell", "Lax-Wendroff", "Beam-Warming", "Fromm", "CHARM", "HCUS", "HQUICK", "Koren",
    "minmod", "Oshker", "ospre", "smart", "Sweby", "UMIST", "van Albada 1", "van Albada 2",
    "van Leer", "monotized central", "superbee", "Barth-Jespersen",
]


def limiterIndex(name):
    """0-based index of a limiter name; unknown names fall back to 0 like solverbase.lua:807-812."""
    try:
        return limiterNames.index(name)
    except ValueError:
        print("!!!! couldn't find limiter %r so falling back on %r !!!!!" % (name, limiterNames[0]))
        return 0


def realBytes(precision):
    """cmdline.float / precision -> sizeof(real) (hydro/app.lua:177-179,892)."""
    return {"double": 8, "float": 4, None: 8}[precision]
