/* TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 *
 * The reference's generated code obtains its field arrays with posix_memalign and never
 * clears them (opesci/regulargrid.py:489-490), yet its stencils read ghost cells that no
 * loop ever writes.  In a fresh process those pages happen to be zero.
 * Linking this file makes that de-facto contract explicit and deterministic: every
 * posix_memalign'd block is zero-filled, whatever the heap history.
 */
#include <errno.h>
#include <stdlib.h>
#include <string.h>

int posix_memalign(void **memptr, size_t alignment, size_t size)
{
    size_t rounded = (size + alignment - 1) / alignment * alignment;
    void *p = aligned_alloc(alignment, rounded ? rounded : alignment);
    if (!p)
        return ENOMEM;
    memset(p, 0, rounded);
    *memptr = p;
    return 0;
}
