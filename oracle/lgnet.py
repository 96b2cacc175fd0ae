"""TEST INFRASTRUCTURE -- CPU restatement of LG-Net's SA_Layer (models/model.py:97-123), dense, small sizes only.
Pinned against the unmodified reference module (tests/golden/make_golden_lgnet.py -> ref_lgnet.npz)."""
import torch
import torch.nn.functional as F


def sa_attention_dense(x_q, x_k, x_v):
    """models/model.py:116-119."""
    energy = torch.bmm(x_q, x_k)
    attention = torch.softmax(energy, dim=-1)
    attention = attention / (1e-9 + attention.sum(dim=1, keepdims=True))
    return torch.bmm(x_v, attention)


def sa_layer(x, sd, eps=1e-5):
    """SA_Layer.forward in eval mode from its state dict `sd` (q_conv and k_conv share one weight, :107)."""
    x_q = F.conv1d(x, sd["q_conv.weight"]).permute(0, 2, 1)
    x_k = F.conv1d(x, sd["k_conv.weight"])
    x_v = F.conv1d(x, sd["v_conv.weight"], sd["v_conv.bias"])
    x_r = sa_attention_dense(x_q, x_k, x_v)
    y = F.conv1d(x - x_r, sd["trans_conv.weight"], sd["trans_conv.bias"])
    y = F.batch_norm(y, sd["after_norm.running_mean"], sd["after_norm.running_var"], sd["after_norm.weight"], sd["after_norm.bias"],
                     training=False, eps=eps)
    return x + torch.relu(y)
