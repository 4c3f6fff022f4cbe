import math


def angle_normalize(x):
    """((x + pi) mod 2 pi) - pi with floored modulo (reference: gops/utils/math_utils.py:8-11)."""
    return ((x + math.pi) % (2 * math.pi)) - math.pi
