"""Model selector with the reference's contract (get_model.py:15-17): name -> module.LSTM_model(**kwargs)."""
import CMPC_model  # noqa: F401


def get_segmentation_model(name, **kwargs):
    if name != "CMPC_model":
        raise NameError("only 'CMPC_model' is provided (the north-star path); got %r" % (name,))
    return CMPC_model.LSTM_model(**kwargs)
