"""The model calls of the reference's three ``tests/visual_*.py`` scripts, without matplotlib: the same data, the same
estimators and hyper-parameters, the same query grids (1001 x 1001 and 500 x 500 points through ``decision_function`` /
``predict`` -- the reference's own large-query call pattern for the kernel mat-vec K7).  Each function returns the arrays
the script would have plotted.  ``models`` is the package to take the estimators from (default: pysvm_b200)."""
from __future__ import annotations

import contextlib
import io

import numpy as np


def _quiet(f):
    with contextlib.redirect_stdout(io.StringIO()):
        return f()


def classify(models=None, grid: int = 1001):
    """tests/visual_classify.py:1-82: LinearSVC and KernelSVC on a 2-feature make_classification problem, decision
    values and labels over a grid x grid mesh."""
    from sklearn.datasets import make_classification
    import pysvm_b200 as P
    models = models or P
    X, y = make_classification(n_samples=250, n_classes=2, n_features=2, n_redundant=0, random_state=2022)
    x_min, x_max = np.min(X[:, 0]), np.max(X[:, 0])
    plot_x = np.linspace(x_min - 1, x_max + 1, grid)
    plot_y = np.linspace(x_min - 1, x_max + 1, grid)          # (sic: the script uses x_min / x_max for both axes)
    xx, yy = np.meshgrid(plot_x, plot_y)
    pts = np.c_[xx.ravel(), yy.ravel()]
    out = {"X": X, "y": y, "grid": pts}
    for name, cls in (("linear", models.LinearSVC), ("kernel", models.KernelSVC)):
        clf = _quiet(lambda: cls().fit(X, y))
        out[name + "_decision"] = np.asarray(clf.decision_function(pts)).reshape(xx.shape)
        out[name + "_labels"] = np.asarray(clf.predict(pts)).reshape(xx.shape)
        out[name + "_model"] = clf
    return out


def regression(models=None):
    """tests/visual_regression.py:1-47: LinearSVR(C=10), NuSVR(poly, degree 2, C=10), KernelSVR(C=10, gamma=0.25)."""
    from sklearn.datasets import make_regression
    import pysvm_b200 as P
    models = models or P
    X, y = make_regression(n_features=1, noise=3, n_samples=50, random_state=2022)
    out = {"X": X, "y": y}
    m = _quiet(lambda: models.LinearSVR(C=10).fit(X, y - 100))
    tx = np.linspace(X.min(0), X.max(0), 2)
    out["linear_x"], out["linear_pred"] = tx, np.asarray(m.predict(tx))
    m = _quiet(lambda: models.NuSVR(kernel='poly', degree=2, C=10).fit(X, 0.01 * y ** 2))
    tx = np.linspace(X.min(0), X.max(0), 100)
    out["nu_x"], out["nu_pred"] = tx, np.asarray(m.predict(tx))
    m = _quiet(lambda: models.KernelSVR(C=10, gamma=0.25).fit(X, 100 * np.sin(0.01 * y) + 100))
    out["rbf_x"], out["rbf_pred"] = tx, np.asarray(m.predict(tx))
    return out


def outlier(models=None, seed: int = 0, grid: int = 500):
    """tests/visual_outlier.py:1-67: OneClassSVM(nu=0.1, gamma=0.1) on two Gaussian blobs, decision values over a
    grid x grid mesh (the script is unseeded; a seed makes the replay comparable)."""
    import pysvm_b200 as P
    models = models or P
    np.random.seed(seed)
    xx, yy = np.meshgrid(np.linspace(-5, 5, grid), np.linspace(-5, 5, grid))
    X = 0.3 * np.random.randn(100, 2)
    X_train = np.r_[X + 2, X - 2]
    X = 0.3 * np.random.randn(20, 2)
    X_test = np.r_[X + 2, X - 2]
    X_outliers = np.random.uniform(low=-4, high=4, size=(20, 2))
    clf = _quiet(lambda: models.OneClassSVM(nu=0.1, kernel="rbf", gamma=0.1).fit(X_train))
    pts = np.c_[xx.ravel(), yy.ravel()]
    return {"X_train": X_train, "grid": pts, "decision": np.asarray(clf.decision_function(pts)).reshape(xx.shape),
            "n_error_train": int((clf.predict(X_train) == -1).sum()), "n_error_test": int((clf.predict(X_test) == -1).sum()),
            "n_error_outliers": int((clf.predict(X_outliers) == 1).sum()), "X_test": X_test, "X_outliers": X_outliers, "model": clf}
