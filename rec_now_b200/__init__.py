"""rec_now_b200 — B200-native (sm_100a) implementation of rec_now's in-batch ranking-loss hot path.

Public surface (same names as the reference's rec_now.rec_block modules):
    rec_now_b200.rec_block.pairwise_loss_from_batch   pairwise_loss, bpr_loss_func, occurance_power_weight, ...
    rec_now_b200.rec_block.listwise_loss_from_batch   to_listwise_sample, listwise_loss_via_softmax_cross_entropy_with_logits, ...
    rec_now_b200.ops                                  thin wrappers over the C ABI (include/recnow_b200.h)
    rec_now_b200.global_mode                          multi-GPU "global in-batch" pairwise loss
The compute lives in librecnow_b200.so (built by `python -m rec_now_b200.build`); there is no CPU fallback.
"""
__version__ = "0.1.0"
