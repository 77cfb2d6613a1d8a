/* bcnn_node.c -- growth of a node's src[] / dst[] index lists
 * (reference src/bcnn_node.c:28-48). */
#include "bcnn_node.h"

#include "bcnn_net.h"
#include "bcnn_utils.h"

static bcnn_status push_index(bcnn_net *net, int **list, int *count, int index) {
    int *grown = (int *)realloc(*list, (size_t)(*count + 1) * sizeof(int));
    BCNN_CHECK_AND_LOG(net->log_ctx, grown != NULL, BCNN_FAILED_ALLOC,
                       "Internal allocation error\n");
    grown[*count] = index;
    *list = grown;
    *count += 1;
    return BCNN_SUCCESS;
}

bcnn_status bcnn_node_add_input(bcnn_net *net, bcnn_node *node, int index) {
    return push_index(net, &node->src, &node->num_src, index);
}

bcnn_status bcnn_node_add_output(bcnn_net *net, bcnn_node *node, int index) {
    return push_index(net, &node->dst, &node->num_dst, index);
}
