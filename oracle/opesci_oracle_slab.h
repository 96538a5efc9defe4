/* TEST INFRASTRUCTURE ONLY -- host-driven halo exchange for the CPU oracle.
 *
 * The CUDA library exchanges slab halos with NCCL.  The oracle has no communication layer: the test
 * installs a callback (e.g. torch.distributed / gloo send+recv) that is called once per field and
 * exchange with the byte ranges to send to / receive from the lower and upper x-neighbour
 * (NULL where the slab end is a physical face).  Geometry: include/opesci_slab.h. */
#ifndef OPESCI_ORACLE_SLAB_H
#define OPESCI_ORACLE_SLAB_H
#include <stddef.h>
typedef int (*opesci_oracle_exchange_fn)(void *user, void *send_lo, void *recv_lo, void *send_hi, void *recv_hi,
                                         size_t nbytes);
void opesci_oracle_set_exchange(opesci_oracle_exchange_fn fn, void *user);
#endif
