from .lgn_tests import covariance_test, lgn_tests, permutation_invariance_test
from .utils import display_err, get_avg_internal_dev, get_avg_output_dev, get_dev, get_node_dev, get_output, plot_all_dev
