/* rkfd_ztk.h - reader for the ZTK model files of the reference (example/model/*.ztk):
 * [roki::chain], [roki::motor], [zeo::shape], [roki::link], [roki::chain::init], [roki::contact].
 * Replaces [EXT] rkChainReadZTK / rkContactInfoArrayReadZTK (call sites reference rkfd_sim.c:229, :264). */
#ifndef RKFD_ZTK_H
#define RKFD_ZTK_H

#include <string>
#include <vector>

#include "rkfd_model.h"

namespace rkfd {
bool ztk_read_chain(const char *filename, ChainHost &chain, std::string &err);
bool ztk_read_contact_info(const char *filename, std::vector<ContactInfoHost> &ci, std::string &err);
}
#endif
