"""Training-mode (autograd) entry points.  Forward/backward kernels for training land after the
inference path; until then training-mode calls fail loudly instead of
silently running some other implementation."""


def _nyi(*a, **k):
    raise NotImplementedError("navc training path (backward kernels) is not built yet; call model.eval() "
                              "or wrap the call in torch.no_grad() for the inference kernels")


encode_train = decoder_forward_train = vocab_forward_train = _nyi
