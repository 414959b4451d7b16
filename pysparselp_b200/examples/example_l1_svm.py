"""L1-regularised multi-class SVM as an LP (reference ``pysparselp/examples/example_l1_svm.py``)."""
import numpy as np

from ..SparseLP import SparseLP, solving_methods


class L1SVM(SparseLP):
    """Zhu, Rosset, Hastie, Tibshirani: 1-norm support vector machines (NIPS 2004)."""

    def add_abs_penalization(self, indices, coef_penalization):
        """aux >= |x[indices]| with cost ``coef_penalization`` (rows ``x - aux <= 0`` then ``-x - aux <= 0``)."""
        aux = self.add_variables_array(indices.size, upper_bounds=None, lower_bounds=0)
        if np.isscalar(coef_penalization):
            assert coef_penalization > 0
        else:
            assert coef_penalization.shape == aux.shape and np.min(coef_penalization) >= 0
        self.set_costs_variables(aux, np.ones(aux.shape) * coef_penalization)
        cols = np.column_stack((indices.ravel(), aux.ravel()))
        for signs in ((1, -1), (-1, -1)):
            self.add_inequality_constraints(cols, np.tile(np.array(signs), (indices.size, 1)),
                                            lower_bounds=None, upper_bounds=0)

    def set_data(self, x, classes, nb_classes=None):
        nb_examples, nb_features = x.shape
        xh = np.hstack((x, np.ones((nb_examples, 1))))
        assert nb_examples == len(classes)
        if nb_classes is None:
            nb_classes = np.max(classes) + 1
        self.weightsIndices = self.add_variables_array((nb_classes, nb_features + 1), None, None)
        self.add_abs_penalization(self.weightsIndices, 1)
        self.epsilonsIndices = self.add_variables_array((nb_examples, 1), upper_bounds=None, lower_bounds=0, costs=1)
        margin = np.ones((nb_examples, nb_classes))
        margin[np.arange(nb_examples), classes] = 0
        own_cols = self.weightsIndices[classes, :]
        for k in range(nb_classes):
            keep = classes != k
            other_cols = np.tile(self.weightsIndices[[k], :], (nb_examples, 1))
            cols = np.column_stack((own_cols, other_cols, self.epsilonsIndices))
            vals = np.column_stack((xh, -xh, np.ones(self.epsilonsIndices.shape)))
            # W[y_i].xh_i - W[k].xh_i + eps_i >= 1   for every example of another class
            self.add_inequality_constraints(cols[keep, :], vals[keep, :], lower_bounds=margin[keep, k],
                                            upper_bounds=None)

    def train(self, method="chambolle_pock_ppd", nb_iter=2000, **solver_options):
        sol, _ = self.solve(method=method, get_timing=True, nb_iter=nb_iter, max_time=np.inf,
                            plot_solution=None, **solver_options)
        self.weights = sol[self.weightsIndices]
        self.activeSet = np.nonzero(sol[self.epsilonsIndices] > 1e-3)[0]

    def classify(self, x):
        xh = np.hstack((x, np.ones((x.shape[0], 1))))
        return np.argmax(xh.dot(self.weights.T), axis=1)


def make_data(nb_examples=1000, nb_classes=3, nb_features=2):
    np.random.seed(1)
    x = np.random.rand(nb_examples, nb_features)
    xh = np.hstack((x, np.ones((nb_examples, 1))))
    weights = np.random.randn(nb_classes, nb_features)
    weights = weights / np.sum(weights ** 2, axis=1)[:, None]
    weights = np.hstack((weights, -0.5 * np.sum(weights, axis=1)[:, None]))
    classes = np.argmax(weights.dot(xh.T).T, axis=1)
    return x, classes


def run(display=False, **solver_options):
    x, classes = make_data()
    svm = L1SVM()
    svm.set_data(x, classes)
    percent_valid = {}
    for method in solving_methods:
        svm.train(method=method, **solver_options)
        percent_valid[method] = 100 * np.mean(classes == svm.classify(x))
    return percent_valid
