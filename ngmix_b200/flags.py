"""
Flag bit values of the fitting results; the values are part of the result
contract and identical to the reference's (ngmix/flags.py:3-28).
"""

NO_ATTEMPT = 2**0
CEN_SHIFT = 2**1
NONPOS_FLUX = 2**2
NONPOS_SIZE = 2**3
LOW_DET = 2**4
MAXITER = 2**5
NONPOS_VAR = 2**6
GMIX_RANGE_ERROR = 2**7
NONPOS_SHAPE_VAR = 2**8
LM_SINGULAR_MATRIX = 2**9
LM_NEG_COV_EIG = 2**10
LM_NEG_COV_DIAG = 2**11
LM_FUNC_NOTFINITE = 2**12
EIG_NOTFINITE = 2**13
DIV_ZERO = 2**14
ZERO_DOF = 2**15

EM_RANGE_ERROR = GMIX_RANGE_ERROR
EM_MAXITER = MAXITER
BAD_VAR = NONPOS_VAR

NAME_MAP = {
    NO_ATTEMPT: 'no attempt',
    CEN_SHIFT: 'center shifted too far',
    NONPOS_FLUX: 'flux <= 0',
    NONPOS_SIZE: 'T <= 0',
    LOW_DET: 'determinant near zero',
    MAXITER: 'max iterations reached',
    NONPOS_VAR: 'non-positive (definite) variance',
    NONPOS_SHAPE_VAR: 'non-positive shape variance',
    GMIX_RANGE_ERROR: 'GMixRangeError raised',
    LM_SINGULAR_MATRIX: 'singular matrix in LM',
    LM_NEG_COV_EIG: 'negative covariance eigenvalue in LM',
    LM_NEG_COV_DIAG: 'negative covariance diagional value in LM',
    LM_FUNC_NOTFINITE: 'function not finite in LM',
    EIG_NOTFINITE: 'eigenvalues of covariance cannot be found in LM',
    DIV_ZERO: 'divide by zero',
    ZERO_DOF: 'degrees of freedom for it is zero (no chi^2/dof possible)',
}


def get_flags_str(val, name_map=None):
    """descriptions of the set bits, joined by '|' (flags.py:55-106)"""
    if name_map is None:
        name_map = NAME_MAP
    if val < 0:
        raise ValueError(f"Flag value {val} must be non-negative.")
    val = int(val) & 0xFFFFFFFF
    names = []
    for power in range(31):
        bit = 1 << power
        if val & bit:
            names.append(name_map.get(bit, "bit 2**%d" % power))
    return "|".join(names)
