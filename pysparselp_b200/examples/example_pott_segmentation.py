"""Potts image-segmentation LP (reference ``pysparselp/examples/example_pott_segmentation.py``).

``ImageLP`` / ``build_linear_program`` / ``run`` keep the reference's names and
argument meaning (``:12-51, :54-92, :95-196``).  The exact ground truth comes from a
max-flow solve; PyMaxflow is optional — when absent, ``scipy.sparse.csgraph.maximum_flow``
on the same grid graph is used (same minimum cut; the label of a pixel is 1 iff it
cannot reach the sink in the residual graph).
"""
import numpy as np

from ..SparseLP import SparseLP, solving_methods


class ImageLP(SparseLP):
    """SparseLP with helpers for pairwise |x_a - x_b| penalties on a pixel grid."""

    def add_penalized_differences(self, ids1, ids2, coef_penalization):
        """One auxiliary variable t >= |x[ids1] - x[ids2]| per pair, cost ``coef_penalization``.

        Row layout (the solver's summation order depends on it): first the block
        ``x1 - x2 - t <= 0`` for all pairs, then the block ``-x1 + x2 - t <= 0``;
        columns inside a row are ``(ids1, ids2, aux)``.
        """
        assert ids1.size == ids2.size
        span = np.maximum(self.upper_bounds[ids1] - self.lower_bounds[ids2],
                          self.upper_bounds[ids2] - self.lower_bounds[ids1])
        aux = self.add_variables_array(ids1.shape, upper_bounds=span, lower_bounds=0, costs=coef_penalization)
        if np.isscalar(coef_penalization):
            assert coef_penalization > 0
        else:
            assert coef_penalization.shape == aux.shape and np.min(coef_penalization) >= 0
        cols = np.column_stack((ids1.ravel(), ids2.ravel(), aux.ravel()))
        for signs in ((1, -1, -1), (-1, 1, -1)):
            vals = np.tile(np.array(signs), (ids1.size, 1))
            self.add_inequality_constraints(cols, vals, lower_bounds=None, upper_bounds=0)

    def add_pott_horizontal(self, indices, coef_penalization):
        self.add_penalized_differences(indices[:, 1:], indices[:, :-1], coef_penalization)

    def add_pott_vertical(self, indices, coef_penalization):
        self.add_penalized_differences(indices[1:, :], indices[:-1, :], coef_penalization)

    def add_pott_model(self, indices, coef_penalization):
        self.add_pott_horizontal(indices, coef_penalization)
        self.add_pott_vertical(indices, coef_penalization)


def graph_cut_labels(unary_terms, pairwise_weight):
    """Exact binary Potts minimiser on a 2-D grid (x in {0,1}, cost unary*x + w*|x_a-x_b|)."""
    try:
        import maxflow

        g = maxflow.Graph[int](0, 0)
        nodeids = g.add_grid_nodes(unary_terms.shape)
        g.add_grid_edges(nodeids, pairwise_weight)
        g.add_grid_tedges(nodeids, unary_terms * 0, unary_terms)
        g.maxflow()
        return np.int_(np.logical_not(g.get_grid_segments(nodeids)))
    except ImportError:
        pass
    import scipy.sparse as sp
    from scipy.sparse.csgraph import breadth_first_order, maximum_flow

    shape = unary_terms.shape
    nn = int(np.prod(shape))
    ids = np.arange(nn).reshape(shape)
    s, t = nn, nn + 1
    rows, cols, caps = [], [], []
    for axis in range(len(shape)):
        a = np.moveaxis(ids, axis, 0)
        u, v = a[:-1].ravel(), a[1:].ravel()
        rows += [u, v]
        cols += [v, u]
        caps += [np.full(u.size, int(pairwise_weight))] * 2
    un = unary_terms.ravel().astype(np.int64)
    src_cap = np.maximum(-un, 0)   # tedge(0, u) with negative u  ==  tedge(-u, 0)
    snk_cap = np.maximum(un, 0)
    rows += [np.full(nn, s), ids.ravel()]
    cols += [ids.ravel(), np.full(nn, t)]
    caps += [src_cap, snk_cap]
    g = sp.csr_matrix((np.concatenate(caps).astype(np.int32), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(nn + 2, nn + 2))
    flow = maximum_flow(g, s, t).flow
    residual = (g - flow).tocsr()
    residual.data = np.maximum(residual.data, 0)
    residual.eliminate_zeros()
    reach_t = np.zeros(nn + 2, dtype=bool)
    reach_t[breadth_first_order(residual.T.tocsr(), t, directed=True, return_predecessors=False)] = True
    return np.int_(np.logical_not(reach_t[:nn].reshape(shape)))


def build_linear_program(image_size, coef_potts, coef_mul, with_ground_truth=True):
    np.random.seed(1)
    size_image = (image_size, image_size, 1)
    unary_terms = np.round(coef_mul * (np.random.rand(*size_image) * 2 - 1))
    coef_potts = round(coef_potts * coef_mul)
    ground_truth = graph_cut_labels(unary_terms, coef_potts) if with_ground_truth else None
    lp = ImageLP()
    indices = lp.add_variables_array(shape=size_image, lower_bounds=0, upper_bounds=1, costs=unary_terms / coef_mul)
    lp.add_pott_model(indices, coef_potts / coef_mul)
    return lp, ground_truth, indices, unary_terms


def run(display=False, image_size=50, coef_mul=500, coef_potts=0.5, max_time=150, nb_iter=100000,
        nb_iter_plot=500, **solver_options):
    """Solve with every available method; returns ``{method: distance_to_ground_truth curve}``."""
    lp, ground_truth, ground_truth_indices, _ = build_linear_program(image_size, coef_potts, coef_mul)
    curves = {}
    for method in solving_methods:
        lp.solve(method=method, get_timing=True, nb_iter=nb_iter, max_time=max_time,
                 ground_truth=ground_truth, ground_truth_indices=ground_truth_indices,
                 plot_solution=None, nb_iter_plot=nb_iter_plot, **solver_options)
        curves[method] = lp.distance_to_ground_truth
    return curves
