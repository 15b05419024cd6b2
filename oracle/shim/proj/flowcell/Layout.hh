// Shadows flowcell/Layout.hh of the reference: the hot path only asks a layout for its read metadata.
#ifndef iSAAC_FLOWCELL_LAYOUT_HH
#define iSAAC_FLOWCELL_LAYOUT_HH
#include <vector>
#include <algorithm>
#include "flowcell/ReadMetadata.hh"
namespace isaac { namespace flowcell {
class Layout {
public:
    Layout() {}
    explicit Layout(const ReadMetadataList &l) : readMetadataList_(l) {}
    const ReadMetadataList &getReadMetadataList() const { return readMetadataList_; }
private:
    ReadMetadataList readMetadataList_;
};
typedef std::vector<Layout> FlowcellLayoutList;
inline unsigned getMaxTotalReadLength(const FlowcellLayoutList &l) { unsigned r = 0; for (const Layout &f : l) r = std::max(r, getTotalReadLength(f.getReadMetadataList())); return r; }
inline unsigned getMaxReadLength(const FlowcellLayoutList &l) { unsigned r = 0; for (const Layout &f : l) for (const ReadMetadata &m : f.getReadMetadataList()) r = std::max(r, m.getLength()); return r; }
} }
#endif
