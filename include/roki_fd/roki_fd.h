/* roki_fd.h - umbrella header, as in the reference (include/roki_fd/roki_fd.h): user programs include
 * only this file.  The whole C-ABI of the B200 build lives in rkfd_b200.h. */
#ifndef ROKI_FD_UMBRELLA_H
#define ROKI_FD_UMBRELLA_H
#include <roki_fd/rkfd_b200.h>
#endif
