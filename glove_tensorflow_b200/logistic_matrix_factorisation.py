"""``python -m glove_tensorflow_b200.logistic_matrix_factorisation`` -- drop-in for the reference's logistic variant
(src/models/logistic_matrix_factorisation.py:65-86): same layer, two weighted sigmoid-CE heads (positives weighted by
``--pos-name``, negatives by ``--neg-name``, combined [1, --neg-factor]); columns overridden to [row, col, pos, neg]."""
import logging

from . import config_utils, estimator

HEAD = "logistic"


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(name)s - %(message)s")
    params = config_utils.parse_args(argv)
    params["input_fn_args"].update({
        "select_columns": [params["row_name"], params["col_name"], params["pos_name"], params["neg_name"]],
        "target_names": [],
    })
    config_utils.save_params(params)
    estimator.train(params, head=HEAD, value_names=(params["pos_name"], params["neg_name"]))


if __name__ == "__main__":
    try:
        main()
    except KeyboardInterrupt:
        pass
