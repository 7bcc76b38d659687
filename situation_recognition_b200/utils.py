"""Helpers API-compatible with the reference's `utils/utils.py`."""
import torch


def load_net(fname, net_list):
    """utils.py:5-31: copy every tensor of checkpoint['model_state_dict'] whose key exists in the model (keys that
    are missing are reported, not fatal).  Raises on a shape mismatch instead of dropping into pdb."""
    for net in net_list:
        device = next(net.parameters()).device
        checkpoint = torch.load(fname, map_location=device, weights_only=False)
        saved = checkpoint['model_state_dict']
        target = getattr(net, "module", net).state_dict()
        with torch.no_grad():
            for k, v in target.items():
                if k in saved:
                    if tuple(saved[k].shape) != tuple(v.shape):
                        raise ValueError('[Error loading] parameter[{}] size mismatch.'.format(k))
                    v.copy_(saved[k])
                else:
                    print('[Missed]: {}'.format(k), v.size())


def format_dict(d, s, p):
    """utils.py:34-42: 'p<key>: <value*100 formatted with s>' joined by ', ' (original imSitu metric format)."""
    return ", ".join(p + str(k) + ": " + s.format(v * 100) for k, v in d.items())
