"""transforms3d.quaternions subset (w, x, y, z convention): mat2quat, quat2mat, qinverse, qmult, rotate_vector.
Used by the reference's submission.py:14 and lib/datasets/mapfree.py:7; written from the standard formulas."""
import numpy as np


def quat2mat(q):
    w, x, y, z = np.asarray(q, dtype=np.float64)
    n = w * w + x * x + y * y + z * z
    if n < np.finfo(np.float64).eps:
        return np.eye(3)
    s = 2.0 / n
    X, Y, Z = x * s, y * s, z * s
    wX, wY, wZ, xX, xY, xZ, yY, yZ, zZ = w * X, w * Y, w * Z, x * X, x * Y, x * Z, y * Y, y * Z, z * Z
    return np.array([[1.0 - (yY + zZ), xY - wZ, xZ + wY],
                     [xY + wZ, 1.0 - (xX + zZ), yZ - wX],
                     [xZ - wY, yZ + wX, 1.0 - (xX + yY)]])


def mat2quat(M):
    """Rotation matrix -> unit quaternion with w >= 0 (largest-eigenvector method, robust to slight non-orthogonality)."""
    Qxx, Qyx, Qzx, Qxy, Qyy, Qzy, Qxz, Qyz, Qzz = np.asarray(M, dtype=np.float64).flat
    K = np.array([[Qxx - Qyy - Qzz, 0, 0, 0],
                  [Qyx + Qxy, Qyy - Qxx - Qzz, 0, 0],
                  [Qzx + Qxz, Qzy + Qyz, Qzz - Qxx - Qyy, 0],
                  [Qyz - Qzy, Qzx - Qxz, Qxy - Qyx, Qxx + Qyy + Qzz]]) / 3.0
    vals, vecs = np.linalg.eigh(K)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    return -q if q[0] < 0 else q


def qconjugate(q):
    return np.asarray(q, dtype=np.float64) * np.array([1.0, -1, -1, -1])


def qinverse(q):
    q = np.asarray(q, dtype=np.float64)
    return qconjugate(q) / np.dot(q, q)


def qmult(q1, q2):
    w1, x1, y1, z1 = q1
    w2, x2, y2, z2 = q2
    return np.array([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                     w1 * y2 + y1 * w2 + z1 * x2 - x1 * z2, w1 * z2 + z1 * w2 + x1 * y2 - y1 * x2])


def rotate_vector(v, q, is_normalized=True):
    varr = np.zeros(4)
    varr[1:] = v
    return qmult(q, qmult(varr, qconjugate(q) if is_normalized else qinverse(q)))[1:]
