// Shadows flowcell/ReadMetadata.hh of the reference (which pulls Boost.Filesystem): same accessors.
#ifndef iSAAC_FLOWCELL_READ_METADATA_HH
#define iSAAC_FLOWCELL_READ_METADATA_HH
#include <vector>
#include <iostream>
namespace isaac { namespace flowcell {
class ReadMetadata {
public:
    ReadMetadata(const unsigned number, const std::vector<unsigned> &cycleList, unsigned index, unsigned offset, unsigned firstReadCycle)
        : number_(number), cycleList_(cycleList), index_(index), offset_(offset), firstReadCycle_(firstReadCycle) {}
    ReadMetadata(unsigned firstCycle, unsigned lastCycle, unsigned index, unsigned offset)
        : number_(index + 1), index_(index), offset_(offset), firstReadCycle_(firstCycle)
    { for (unsigned c = firstCycle; lastCycle >= c; ++c) cycleList_.push_back(c); }
    virtual ~ReadMetadata() {}
    unsigned getLength() const { return cycleList_.size(); }
    unsigned getFirstReadCycle() const { return firstReadCycle_; }
    unsigned getFirstCycle() const { return cycleList_.front(); }
    unsigned getLastCycle() const { return cycleList_.back(); }
    const std::vector<unsigned> &getCycles() const { return cycleList_; }
    unsigned getIndex() const { return index_; }
    unsigned getNumber() const { return number_; }
    unsigned getOffset() const { return offset_; }
private:
    unsigned number_; std::vector<unsigned> cycleList_; unsigned index_; unsigned offset_; unsigned firstReadCycle_;
};
class ReadMetadataList : public std::vector<ReadMetadata> {
public:
    ReadMetadataList(const std::vector<ReadMetadata> &that) : std::vector<ReadMetadata>(that) {}
    ReadMetadataList() {}
};
inline unsigned getTotalReadLength(const ReadMetadataList &l) { unsigned r = 0; for (const ReadMetadata &m : l) r += m.getLength(); return r; }
inline std::ostream &operator<<(std::ostream &os, const ReadMetadata &r) { return os << "ReadMetadata(" << r.getNumber() << "," << r.getLength() << ")"; }
} }
#endif
