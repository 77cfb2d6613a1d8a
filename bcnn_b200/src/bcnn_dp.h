/*
 * bcnn_dp.h -- data-parallel hooks of the train loop (new: jnbraun/bcnn has no
 * multi-GPU support; the insertion points are its bcnn_backward / bcnn_update loops,
 * src/bcnn_net.c:424-429 and src/bcnn_learner.c:167-175).
 */
#ifndef BCNN_DP_H
#define BCNN_DP_H

#include "bcnn_net.h"

#ifdef __cplusplus
extern "C" {
#endif

int bcnn_dp_world_size(bcnn_net *net);                              /* 1 without DP */
void bcnn_dp_after_node_backward(bcnn_net *net, bcnn_node *node);   /* queue all-reduce */
void bcnn_dp_before_update(bcnn_net *net);                          /* join comm stream */
void bcnn_dp_set_deferred(bcnn_net *net, int on);                   /* backward issues no transfers */
void bcnn_dp_allreduce_all(bcnn_net *net);                          /* all buckets, now */
void bcnn_dp_allreduce_range(bcnn_net *net, int first, int end);    /* nodes [first, end) */
int bcnn_dp_backward_split(bcnn_net *net);
void bcnn_dp_sync(bcnn_net *net);
void bcnn_dp_release(bcnn_net *net);

#ifdef __cplusplus
}
#endif
#endif /* BCNN_DP_H */
