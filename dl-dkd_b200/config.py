"""Model / option defaults for the host mirror (the fields DLDKD(config, opt) and the eval entry points read).

Values follow the reference's training script and option parser: model_config method/train.py:300-314
(+ label_style, read by forward() at method/model.py:138), option defaults method/config.py (hidden sizes :70-71,
eval batch sizes :48-49, loss weights)."""


class AttrDict(dict):
    """dict with attribute access; missing keys raise AttributeError (what the model's getattr probes expect)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def model_config(visual_input_size, query_input_size, hidden=384, n_heads=4, max_ctx_l=128, max_desc_l=30, **over):
    cfg = AttrDict(visual_input_size=visual_input_size, query_input_size=query_input_size,
                   inheritance_hidden=hidden, exploration_hidden=hidden, max_ctx_l=max_ctx_l, max_desc_l=max_desc_l,
                   input_drop=0.2, drop=0.2, n_heads=n_heads, initializer_range=0.02, margin=0.1,
                   use_hard_negative=False, hard_pool_size=20, label_style="soft")
    cfg.update(over)
    return cfg


def options(device="cpu", eval_query_bsz=50, eval_context_bsz=200, **over):
    opt = AttrDict(double_branch=True, kl_intra_weight=0.1, inher_nce_weight=0.04, explore_nce_weight=0.04,
                   collection="tvr", alpha=0.8, belta=0.8, eval_context_bsz=eval_context_bsz,
                   eval_query_bsz=eval_query_bsz, num_workers=0, pin_memory=False, device=device)
    opt.update(over)
    return opt
