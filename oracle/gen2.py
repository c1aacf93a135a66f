"""Oracle for the Gen-2 ("sklearn PCA") path -- test infrastructure only.

Reference call sites:
  * MultiFaceTrainer.train_pca_model            train-v5.py:349-385 (= train-v4.py:110-146)     (F2)
  * MultiModelFaceScanner.extract_face_features scan-template-v4.py:253-268                     (N1, J1)
  * .recognize_face_with_model                  scan-template-v4.py:270-287                     (M1)
  * .recognize_face_all_models                  scan-template-v4.py:289-319                     (M3)
  * ManualPCA / ManualStandardScaler            scripts/manual/train-v2.py:9-72                 (F3)
The arithmetic itself lives in scikit-learn (un-vendored; author's pickle says 1.7.1, container has
1.9.0): StandardScaler (preprocessing/_data.py), PCA._fit_full / _transform (decomposition/_pca.py,
_base.py), svd_flip (utils/extmath.py), cosine_similarity (metrics/pairwise.py).  Their published
algorithms are restated below in numpy and pinned by tests/golden/gen2_joseph.npz, which holds the
outputs of the reference's own train_pca_model / extract_face_features / recognize_face_with_model
run in the build container (tests/golden/make_golden.py).
"""
import numpy as np


# ------------------------------------------------------------------ StandardScaler (N1)
def scaler_fit(X):
    """sklearn StandardScaler.fit (dense path: _incremental_mean_and_var, first batch).

    mean_ = sum/N; var_ = (sum((X-T)^2) - (sum(X-T))^2/N)/N with T = sum/N (population variance);
    scale_ = sqrt(var_) with near-constant features (_is_constant_feature) mapped to 1.0.
    """
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    new_sum = np.sum(X, axis=0)
    mean = new_sum / n
    temp = X - mean
    correction = np.sum(temp, axis=0)
    temp **= 2
    unnorm = np.sum(temp, axis=0)
    unnorm -= correction ** 2 / n
    var = unnorm / n
    eps = np.finfo(np.float64).eps
    constant = var <= n * eps * var + (n * mean * eps) ** 2
    scale = np.sqrt(var)
    scale = np.where(constant | (scale == 0.0), 1.0, scale)
    return mean, var, scale


def scaler_transform(X, mean, scale):
    """StandardScaler.transform: X -= mean_; X /= scale_ (two roundings)."""
    Z = np.asarray(X, dtype=np.float64).copy()
    Z -= mean
    Z /= scale
    return Z


# ------------------------------------------------------------------ PCA full solver (F2)
def svd_flip_v(U, Vt):
    """sklearn svd_flip(u_based_decision=False): largest-|entry| of every Vt row made positive."""
    max_abs = np.argmax(np.abs(Vt), axis=1)
    signs = np.sign(Vt[np.arange(Vt.shape[0]), max_abs])
    signs = np.where(signs == 0, 1.0, signs)
    return U * signs[None, :], Vt * signs[:, None]


def pca_fit_full(Z, n_components):
    """sklearn PCA(n_components)._fit_full + fit_transform on an already standardised matrix.

    Returns dict(components, mean, explained_variance, explained_variance_ratio, singular_values,
                 noise_variance, features) with features = U[:, :k] * S[:k].
    """
    Z = np.asarray(Z, dtype=np.float64)
    n_samples, n_features = Z.shape
    mean = Z.mean(axis=0)
    Zc = Z - mean
    U, S, Vt = np.linalg.svd(Zc, full_matrices=False)
    U, Vt = svd_flip_v(U, Vt)
    explained_variance = (S ** 2) / (n_samples - 1)
    total_var = explained_variance.sum()
    ratio = explained_variance / total_var
    k = n_components
    if k < min(n_features, n_samples):
        noise_variance = explained_variance[k:].mean()
    else:
        noise_variance = 0.0
    return {
        'components': Vt[:k],
        'mean': mean,
        'explained_variance': explained_variance[:k],
        'explained_variance_ratio': ratio[:k],
        'singular_values': S[:k],
        'noise_variance': noise_variance,
        'features': U[:, :k] * S[:k],
    }


def pca_transform(Z, components, mean):
    """sklearn _BasePCA.transform (whiten=False): X @ components_.T - mean_ @ components_.T."""
    Z = np.asarray(Z, dtype=np.float64)
    return Z @ components.T - (mean.reshape(1, -1) @ components.T)


def train_pca_model(face_images, n_components):
    """train-v5.py:349-385 on face_images uint8 [N, D].  Returns a dict of everything the trainer stores."""
    X = np.asarray(face_images)
    mean_face = np.mean(X, axis=0)                                        # :366
    s_mean, s_var, s_scale = scaler_fit(X)                                # :370
    Z = scaler_transform(X, s_mean, s_scale)
    fit = pca_fit_full(Z, n_components)                                   # :373
    return {
        'mean_face': mean_face,
        'scaler_mean': s_mean, 'scaler_var': s_var, 'scaler_scale': s_scale,
        'eigenfaces': fit['components'],                                  # :376
        'face_features': fit['features'],                                 # :382
        'pca_mean': fit['mean'],
        'explained_variance': fit['explained_variance'],
        'explained_variance_ratio': fit['explained_variance_ratio'],
        'singular_values': fit['singular_values'],
        'noise_variance': fit['noise_variance'],
    }


# ------------------------------------------------------------------ recognition (J1, M1, M3)
def extract_features(flat_u8, scaler_mean, scaler_scale, components, pca_mean):
    """scan-template-v4.py:263-268 for one or many flattened crops [B, D] -> float64 [B, k]."""
    X = np.atleast_2d(np.asarray(flat_u8))
    Z = scaler_transform(X, scaler_mean, scaler_scale)
    return pca_transform(Z, components, pca_mean)


def sk_normalize_rows(A):
    """sklearn.preprocessing.normalize(norm='l2'): rows / sqrt(einsum('ij,ij->i')), zero norm -> 1."""
    A = np.asarray(A, dtype=np.float64)
    norms = np.sqrt(np.einsum('ij,ij->i', A, A))
    norms = np.where(norms == 0.0, 1.0, norms)
    return A / norms[:, None]


def sk_cosine_similarity(X, Y):
    """sklearn.metrics.pairwise.cosine_similarity: normalise both sides, then X_n @ Y_n.T."""
    return sk_normalize_rows(X) @ sk_normalize_rows(Y).T


def recognize_with_model(face_features, gallery, face_labels, person_id_map, threshold=0.7):
    """scan-template-v4.py:270-287.  Returns (person_id, person_name, max_similarity)."""
    sims = sk_cosine_similarity([face_features], gallery)[0]              # :274
    max_idx = int(np.argmax(sims))                                        # :275
    max_similarity = sims[max_idx]
    if max_similarity >= threshold:                                       # :278
        person_id = face_labels[max_idx]
        person_name = "unknown"
        for name, pid in person_id_map.items():                           # :281-284
            if pid == person_id:
                person_name = name
                break
        return person_id, person_name, max_similarity
    return -1, "unknown", max_similarity                                  # :287


def recognize_all_models(flat_u8, models, threshold=0.8):
    """scan-template-v4.py:289-319 on an already preprocessed crop.

    models: ordered dict person_name -> dict(scaler_mean, scaler_scale, components, pca_mean,
            face_features, face_labels, person_id_map).
    """
    best_result = None
    best_confidence = 0.0
    for person_name, m in models.items():
        feats = extract_features(flat_u8, m['scaler_mean'], m['scaler_scale'], m['components'], m['pca_mean'])[0]
        pid, name, conf = recognize_with_model(feats, m['face_features'], m['face_labels'],
                                               m['person_id_map'], threshold)
        if conf > best_confidence:                                        # :306 (strict)
            best_confidence = conf
            best_person = name if name != "unknown" else person_name      # :308
            best_result = (pid, best_person, conf)
    if best_result:
        return best_result
    return -1, "unknown", 0.0


def recognize_batch(flat_u8, m, threshold=0.7):
    """Batched extract_features + cosine argmax.  Returns (best[B], argmax[B], label_or_-1[B])."""
    feats = extract_features(flat_u8, m['scaler_mean'], m['scaler_scale'], m['components'], m['pca_mean'])
    sims = sk_cosine_similarity(feats, m['face_features'])
    idx = np.argmax(sims, axis=1)
    best = sims[np.arange(len(idx)), idx]
    labels = np.where(best >= threshold, np.asarray(m['face_labels'])[idx], -1)
    return best, idx, labels


# ------------------------------------------------------------------ scripts/manual variant (F3)
def manual_scaler_fit(X):
    """ManualStandardScaler.fit, scripts/manual/train-v2.py:53-63: std == 0 -> 1."""
    X = np.asarray(X, dtype=np.float64)
    mean = np.mean(X, axis=0)
    std = np.std(X, axis=0)
    std = np.where(std == 0, 1.0, std)
    return mean, std


def manual_pca_fit(Z, n_components):
    """ManualPCA.fit, scripts/manual/train-v2.py:9-45: D x D np.cov + eigh, top-k rows."""
    Z = np.asarray(Z, dtype=np.float64)
    mean = np.mean(Z, axis=0)
    Zc = Z - mean
    cov = np.cov(Zc.T)
    w, v = np.linalg.eigh(cov)
    idx = np.argsort(w)[::-1]
    w, v = w[idx], v[:, idx]
    return {'mean': mean, 'components': v[:, :n_components].T,
            'explained_variance': w[:n_components],
            'explained_variance_ratio': w[:n_components] / np.sum(w)}
