// Minimal libfive-compatible archive reader (host only): the read side of
// libfive's Archive (include/libfive/tree/archive.hpp:18-60), enough for
// `libfive::Archive::deserialize(ifs).shapes.front().tree` as used by every
// reference benchmark driver (e.g. benchmark/render_2d_table.cpp:34-35).
// Format: libfive/libfive/src/tree/deserializer.cpp:38-143.
#pragma once
#include <iosfwd>
#include <list>
#include <map>
#include <string>

#include "libfive/tree/tree.hpp"

namespace libfive {

class Archive {
public:
    struct Shape {
        Tree tree = Tree::Invalid();
        std::string name;
        std::string doc;
        std::map<Tree::Id, std::string> vars;
    };
    std::list<Shape> shapes;

    static Archive deserialize(std::istream& in);
};

}  // namespace libfive
