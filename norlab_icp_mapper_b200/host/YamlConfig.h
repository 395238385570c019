// YamlConfig.h -- Mapper::loadYamlConfig (reference Mapper.cpp:59-185) without yaml-cpp: a reader for the YAML subset the
// reference's configuration files use (block maps and lists by indentation, `- Name:` list entries with a nested parameter map,
// scalars, `[a, b]` / `{k: v}` flow collections of scalars, `#` comments) and the mapping of the libpointmatcher / norlab_icp_mapper names onto MapperConfig.
// /root/reference/examples/config.yaml and the snippets of docs/MapperConfiguration.md load unmodified.  Names this path does not
// implement are refused with InvalidParameter -- what PM::Registrar does for an unknown class -- never ignored.
#pragma once
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "Mapper.h"

namespace norlab_icp_mapper_b200 {
namespace yaml {

struct Node {
    enum Kind { Null, Scalar, Map, List } kind = Null;
    std::string scalar;
    std::vector<std::pair<std::string, Node>> map;  // insertion order kept (checker order matters)
    std::vector<Node> list;
    const Node* find(const std::string& key) const {
        for (const auto& kv : map)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool isNull() const { return kind == Null; }
};

namespace detail {
struct Line {
    int indent;
    std::string text;
    int number;
};
inline std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
    return s.substr(a, b - a);
}
inline std::string unquote(const std::string& s) {
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
    return s;
}
inline Node scalarOrFlow(const std::string& raw) {
    Node n;
    const std::string v = trim(raw);
    if (v.empty() || v == "~" || v == "null") return n;
    if (v.front() == '[' && v.back() == ']') {
        n.kind = Node::List;
        std::string item;
        std::istringstream is(v.substr(1, v.size() - 2));
        while (std::getline(is, item, ',')) {
            Node e;
            e.kind = Node::Scalar;
            e.scalar = unquote(trim(item));
            if (!e.scalar.empty()) n.list.push_back(e);
        }
        return n;
    }
    if (v.front() == '{' && v.back() == '}') {  // flow map of scalars: {type: distance, value: 2.5}
        n.kind = Node::Map;
        std::string item;
        std::istringstream is(v.substr(1, v.size() - 2));
        while (std::getline(is, item, ',')) {
            const size_t c = item.find(':');
            if (c == std::string::npos) throw InvalidParameter("yaml: expected `key: value` inside { }");
            Node e;
            e.kind = Node::Scalar;
            e.scalar = unquote(trim(item.substr(c + 1)));
            n.map.emplace_back(unquote(trim(item.substr(0, c))), e);
        }
        return n;
    }
    n.kind = Node::Scalar;
    n.scalar = unquote(v);
    return n;
}
inline size_t keyColon(const std::string& t) {  // position of the ':' that ends a key, or npos
    for (size_t i = 0; i < t.size(); ++i) {
        if (t[i] == '[' || t[i] == '{' || t[i] == '"' || t[i] == '\'') return std::string::npos;
        if (t[i] == ':' && (i + 1 == t.size() || t[i + 1] == ' ')) return i;
    }
    return std::string::npos;
}
inline Node parseBlock(std::vector<Line>& L, size_t& i, int indent);
inline Node parseValueAfterKey(std::vector<Line>& L, size_t& i, int keyIndent, const std::string& rest) {
    if (!trim(rest).empty()) return scalarOrFlow(rest);
    // nested block (deeper indentation), or a list at the SAME indentation as the key (yaml allows `key:\n- item`)
    if (i < L.size() && (L[i].indent > keyIndent || (L[i].indent == keyIndent && L[i].text.rfind("- ", 0) == 0))) return parseBlock(L, i, L[i].indent);
    return Node();
}
inline Node parseBlock(std::vector<Line>& L, size_t& i, int indent) {
    Node n;
    while (i < L.size() && L[i].indent == indent) {
        const std::string t = L[i].text;
        if (t.rfind("- ", 0) == 0 || t == "-") {
            if (n.kind == Node::Map) throw InvalidParameter("yaml line " + std::to_string(L[i].number) + ": list entry inside a map");
            n.kind = Node::List;
            const std::string rest = t.size() > 1 ? trim(t.substr(2)) : std::string();
            const int inner = indent + 2;  // the entry's own content starts two columns further in
            ++i;
            const size_t c = keyColon(rest);
            if (rest.empty()) {
                n.list.push_back(i < L.size() && L[i].indent > indent ? parseBlock(L, i, L[i].indent) : Node());
            } else if (c == std::string::npos) {
                n.list.push_back(scalarOrFlow(rest));
            } else {  // `- Key: value` / `- Key:` followed by the entry's map: the first pair of a map that may continue below
                Node m;
                m.kind = Node::Map;
                const std::string key = unquote(trim(rest.substr(0, c)));
                Node v = parseValueAfterKey(L, i, inner, rest.substr(c + 1));
                m.map.emplace_back(key, v);
                if (i < L.size() && L[i].indent == inner && L[i].text.rfind("- ", 0) != 0) {
                    Node more = parseBlock(L, i, inner);
                    for (auto& kv : more.map) m.map.push_back(kv);
                }
                n.list.push_back(m);
            }
        } else {
            const size_t c = keyColon(t);
            if (c == std::string::npos) throw InvalidParameter("yaml line " + std::to_string(L[i].number) + ": expected `key: value`");
            if (n.kind == Node::List) throw InvalidParameter("yaml line " + std::to_string(L[i].number) + ": map entry inside a list");
            n.kind = Node::Map;
            const std::string key = unquote(trim(t.substr(0, c)));
            ++i;
            n.map.emplace_back(key, parseValueAfterKey(L, i, indent, t.substr(c + 1)));
        }
    }
    if (i < L.size() && L[i].indent > indent) throw InvalidParameter("yaml line " + std::to_string(L[i].number) + ": unexpected indentation");
    return n;
}
}  // namespace detail

inline Node parse(std::istream& in) {
    std::vector<detail::Line> L;
    std::string raw;
    int number = 0;
    while (std::getline(in, raw)) {
        ++number;
        // strip comments (a '#' at the line start or after a blank, outside quotes)
        bool q1 = false, q2 = false;
        for (size_t k = 0; k < raw.size(); ++k) {
            if (raw[k] == '\'' && !q2) q1 = !q1;
            if (raw[k] == '"' && !q1) q2 = !q2;
            if (raw[k] == '#' && !q1 && !q2 && (k == 0 || raw[k - 1] == ' ' || raw[k - 1] == '\t')) {
                raw.resize(k);
                break;
            }
        }
        if (detail::trim(raw).empty() || detail::trim(raw) == "---") continue;
        int indent = 0;
        while (indent < (int)raw.size() && raw[indent] == ' ') ++indent;
        if (indent < (int)raw.size() && raw[indent] == '\t') throw InvalidParameter("yaml line " + std::to_string(number) + ": tabs are not allowed for indentation");
        L.push_back({indent, detail::trim(raw), number});
    }
    size_t i = 0;
    if (L.empty()) return Node();
    Node root = detail::parseBlock(L, i, L[0].indent);
    if (i != L.size()) throw InvalidParameter("yaml line " + std::to_string(L[i].number) + ": inconsistent indentation");
    return root;
}

}  // namespace yaml

namespace detail_cfg {
inline float num(const yaml::Node& n, const std::string& what) {
    if (n.kind != yaml::Node::Scalar) throw InvalidParameter(what + ": expected a number");
    if (n.scalar == "inf" || n.scalar == ".inf" || n.scalar == "+inf") return std::numeric_limits<float>::infinity();
    char* end = nullptr;
    const double v = std::strtod(n.scalar.c_str(), &end);
    if (end == n.scalar.c_str() || *end != '\0') throw InvalidParameter(what + ": `" + n.scalar + "` is not a number");
    return (float)v;
}
// one `- Name: {params}` / `Name: {params}` entry -> (name, params as strings); unknown parameter names are the callee's business
inline std::pair<std::string, Parameters> namedEntry(const yaml::Node& n, const std::string& where) {
    if (n.kind == yaml::Node::Scalar) return {n.scalar, Parameters()};
    if (n.kind != yaml::Node::Map || n.map.size() != 1) throw InvalidParameter(where + ": expected `Name:` followed by its parameters");
    Parameters p;
    const yaml::Node& body = n.map[0].second;
    if (body.kind == yaml::Node::Map) {
        for (const auto& kv : body.map) {
            if (kv.second.kind == yaml::Node::List) {
                std::string joined;
                for (const auto& e : kv.second.list) joined += (joined.empty() ? "" : ",") + e.scalar;
                p[kv.first] = joined;
            } else {
                p[kv.first] = kv.second.scalar;
            }
        }
    } else if (!body.isNull()) {
        throw InvalidParameter(where + ": the parameters of " + n.map[0].first + " must be a map");
    }
    return {n.map[0].first, p};
}
inline float take(Parameters& p, const std::string& key, float def) {
    auto it = p.find(key);
    if (it == p.end()) return def;
    yaml::Node n;
    n.kind = yaml::Node::Scalar;
    n.scalar = it->second;
    p.erase(it);
    return num(n, key);
}
inline std::string takeStr(Parameters& p, const std::string& key, const std::string& def) {
    auto it = p.find(key);
    if (it == p.end()) return def;
    const std::string v = it->second;
    p.erase(it);
    return v;
}
inline void noneLeft(const Parameters& p, const std::string& owner) {  // PM::Parametrizable: "Parameter X for module Y was set but is not used"
    if (!p.empty()) throw InvalidParameter("Parameter " + p.begin()->first + " for module " + owner + " was set but is not used");
}
}  // namespace detail_cfg

//! Mapper::loadYamlConfig (Mapper.cpp:59-185): the four top-level nodes `input`, `icp`, `post`, `mapper`, each optional.
inline MapperConfig loadYamlConfig(std::istream& in, bool is3D) {
    using namespace detail_cfg;
    const yaml::Node root = yaml::parse(in);
    MapperConfig cfg;
    const int dim = is3D ? 3 : 2;
    b200icp_config_default(&cfg.icp, dim);  // icp.setDefault() when the node is absent (Mapper.cpp:75-78)
    if (!root.isNull() && root.kind != yaml::Node::Map) throw InvalidParameter("the configuration must be a map with the keys input / icp / post / mapper");

    if (const yaml::Node* input = root.find("input")) {  // Mapper.cpp:80-88
        if (!input->isNull() && input->kind != yaml::Node::List) throw InvalidParameter("input: expected a list of DataPointsFilters");
        for (const auto& e : input->list) {
            auto ne = namedEntry(e, "input");
            Parameters& p = ne.second;
            b200icp_filter f{};
            if (ne.first == "BoundingBoxDataPointsFilter") {
                f.kind = B200ICP_FILTER_BOUNDING_BOX;
                f.lo[0] = take(p, "xMin", -1.f); f.hi[0] = take(p, "xMax", 1.f);
                f.lo[1] = take(p, "yMin", -1.f); f.hi[1] = take(p, "yMax", 1.f);
                f.lo[2] = take(p, "zMin", -1.f); f.hi[2] = take(p, "zMax", 1.f);
                f.remove_inside = take(p, "removeInside", 1.f) != 0.f;
                cfg.inputFilters.push_back(f);
            } else if (ne.first == "DistanceLimitDataPointsFilter") {
                f.kind = B200ICP_FILTER_DISTANCE_LIMIT;
                f.dim = (int)take(p, "dim", -1.f);
                f.dist = take(p, "dist", 1.f);
                f.remove_inside = take(p, "removeInside", 1.f) != 0.f;
                cfg.inputFilters.push_back(f);
            } else if (ne.first == "RandomSamplingDataPointsFilter") {
                f.kind = B200ICP_FILTER_RANDOM_SAMPLING;
                f.dist = take(p, "prob", 0.75f);
                f.dim = (int)take(p, "seed", 0.f);
                take(p, "randomSamplingMethod", 0.f);
                cfg.inputFilters.push_back(f);
            } else if (ne.first == "AddDescriptorDataPointsFilter") {
                const std::string name = takeStr(p, "descriptorName", "");
                const int d = (int)take(p, "descriptorDimension", 1.f);
                if (name != "probabilityDynamic" || d != 1)
                    throw InvalidParameter("AddDescriptorDataPointsFilter: only {descriptorName: probabilityDynamic, descriptorDimension: 1} is implemented on this path");
                cfg.addProbabilityDynamic = true;
                cfg.probabilityDynamicValue = take(p, "descriptorValues", 0.6f);
            } else if (ne.first == "SurfaceNormalDataPointsFilter") {
                cfg.inputSurfaceNormalKnn = (int)take(p, "knn", 5.f);
                take(p, "epsilon", 0.f);  // (the search is exact)
                take(p, "keepNormals", 1.f);
            } else if (ne.first == "IdentityDataPointsFilter") {
            } else {
                throw InvalidParameter("Trying to instantiate unknown DataPointsFilter " + ne.first + " (input chain of the B200 path: BoundingBox, DistanceLimit, RandomSampling, AddDescriptor, SurfaceNormal)");
            }
            noneLeft(p, ne.first);
        }
    }

    if (const yaml::Node* icp = root.find("icp")) {  // Mapper.cpp:70-78 -> PM::ICPSequence::loadFromYamlNode
        if (icp->kind != yaml::Node::Map) throw InvalidParameter("icp: expected a map");
        b200icp_config& c = cfg.icp;
        // loadFromYaml starts from an EMPTY chain: no outlier filter, no checker, unless listed
        c.n_outlier = 0;
        c.max_iteration_count = 0;
        c.use_differential = 0;
        c.use_bound = 0;
        for (const auto& kv : icp->map) {
            const std::string& key = kv.first;
            const yaml::Node& v = kv.second;
            if (key == "matcher") {
                auto ne = namedEntry(v, "icp.matcher");
                if (ne.first != "KDTreeMatcher") throw InvalidParameter("Trying to instantiate unknown Matcher " + ne.first);
                c.knn = (int)take(ne.second, "knn", 1.f);
                c.max_dist = take(ne.second, "maxDist", std::numeric_limits<float>::infinity());
                c.epsilon = take(ne.second, "epsilon", 0.f);
                take(ne.second, "searchType", 1.f);
                noneLeft(ne.second, ne.first);
            } else if (key == "outlierFilters") {
                if (!v.isNull() && v.kind != yaml::Node::List) throw InvalidParameter("icp.outlierFilters: expected a list");
                for (const auto& e : v.list) {
                    auto ne = namedEntry(e, "icp.outlierFilters");
                    Parameters& p = ne.second;
                    if (c.n_outlier >= B200ICP_MAX_OUTLIER_FILTERS) throw InvalidParameter("icp.outlierFilters: at most 4 filters");
                    const int i = c.n_outlier++;
                    c.outlier_param2[i] = c.outlier_param3[i] = 0.f;
                    c.outlier_mode[i] = 0;
                    if (ne.first == "TrimmedDistOutlierFilter") {
                        c.outlier_kind[i] = B200ICP_OUTLIER_TRIMMED_DIST;
                        c.outlier_param[i] = take(p, "ratio", 0.85f);
                    } else if (ne.first == "MaxDistOutlierFilter") {
                        c.outlier_kind[i] = B200ICP_OUTLIER_MAX_DIST;
                        c.outlier_param[i] = take(p, "maxDist", 1.f);
                    } else if (ne.first == "MinDistOutlierFilter") {
                        c.outlier_kind[i] = B200ICP_OUTLIER_MIN_DIST;
                        c.outlier_param[i] = take(p, "minDist", 1.f);
                    } else if (ne.first == "MedianDistOutlierFilter") {
                        c.outlier_kind[i] = B200ICP_OUTLIER_MEDIAN_DIST;
                        c.outlier_param[i] = take(p, "factor", 3.f);
                    } else if (ne.first == "VarTrimmedDistOutlierFilter") {
                        c.outlier_kind[i] = B200ICP_OUTLIER_VAR_TRIMMED_DIST;
                        c.outlier_param[i] = take(p, "minRatio", 0.05f);
                        c.outlier_param2[i] = take(p, "maxRatio", 0.99f);
                        c.outlier_param3[i] = take(p, "lambda", 2.35f);
                    } else if (ne.first == "SurfaceNormalOutlierFilter") {
                        c.outlier_kind[i] = B200ICP_OUTLIER_SURFACE_NORMAL;
                        c.outlier_param[i] = take(p, "maxAngle", 1.57f);
                    } else if (ne.first == "RobustOutlierFilter") {
                        c.outlier_kind[i] = B200ICP_OUTLIER_ROBUST;
                        static const char* fcts[] = {"cauchy", "welsch", "sc", "gm", "tukey", "huber", "L1", "student"};
                        static const char* scales[] = {"none", "mad", "berg", "std"};
                        const std::string fct = takeStr(p, "robustFct", "cauchy"), sc = takeStr(p, "scaleEstimator", "mad"),
                                          dt = takeStr(p, "distanceType", "point2point");
                        int fi = -1, si = -1;
                        for (int k = 0; k < 8; ++k)
                            if (fct == fcts[k]) fi = k;
                        for (int k = 0; k < 4; ++k)
                            if (sc == scales[k]) si = k;
                        if (fi < 0) throw InvalidParameter("RobustOutlierFilter: unknown robustFct " + fct);
                        if (si < 0) throw InvalidParameter("RobustOutlierFilter: unknown scaleEstimator " + sc);
                        if (dt != "point2point" && dt != "point2plane") throw InvalidParameter("RobustOutlierFilter: unknown distanceType " + dt);
                        c.outlier_param[i] = take(p, "tuning", 1.f);
                        c.outlier_param2[i] = take(p, "approximation", std::numeric_limits<float>::infinity());
                        c.outlier_mode[i] = B200ICP_ROBUST_MODE(fi, si, dt == "point2plane" ? 1 : 0, (int)take(p, "nbIterationForScale", 0.f));
                    } else if (ne.first == "NullOutlierFilter") {
                        --c.n_outlier;
                    } else {
                        throw InvalidParameter("Trying to instantiate unknown OutlierFilter " + ne.first);
                    }
                    noneLeft(p, ne.first);
                }
            } else if (key == "errorMinimizer") {
                auto ne = namedEntry(v, "icp.errorMinimizer");
                c.minimizer_flags = 0;
                if (ne.first == "PointToPlaneErrorMinimizer") {
                    c.minimizer = B200ICP_MIN_POINT_TO_PLANE;
                    if (take(ne.second, "force2D", 0.f) != 0.f) c.minimizer_flags |= 1;
                    if (take(ne.second, "force4DOF", 0.f) != 0.f) c.minimizer_flags |= 2;
                } else if (ne.first == "PointToPointErrorMinimizer") {
                    c.minimizer = B200ICP_MIN_POINT_TO_POINT;
                } else if (ne.first == "IdentityErrorMinimizer") {
                    c.minimizer = B200ICP_MIN_IDENTITY;
                } else {
                    throw InvalidParameter("Trying to instantiate unknown ErrorMinimizer " + ne.first);
                }
                noneLeft(ne.second, ne.first);
            } else if (key == "transformationCheckers") {
                if (!v.isNull() && v.kind != yaml::Node::List) throw InvalidParameter("icp.transformationCheckers: expected a list");
                c.checker_order = 0;
                bool counterSeen = false;
                for (const auto& e : v.list) {
                    auto ne = namedEntry(e, "icp.transformationCheckers");
                    Parameters& p = ne.second;
                    if (ne.first == "CounterTransformationChecker") {
                        c.max_iteration_count = (int)take(p, "maxIterationCount", 40.f);
                        counterSeen = true;
                    } else if (ne.first == "DifferentialTransformationChecker") {
                        c.use_differential = 1;
                        c.min_diff_rot_err = take(p, "minDiffRotErr", 0.001f);
                        c.min_diff_trans_err = take(p, "minDiffTransErr", 0.001f);
                        c.smooth_length = (int)take(p, "smoothLength", 3.f);
                        if (!counterSeen) c.checker_order |= 1;  // listed before the Counter (see b200icp_config::checker_order)
                    } else if (ne.first == "BoundTransformationChecker") {
                        c.use_bound = 1;
                        c.max_rotation_norm = take(p, "maxRotationNorm", 1.f);
                        c.max_translation_norm = take(p, "maxTranslationNorm", 1.f);
                        if (!counterSeen) c.checker_order |= 2;
                    } else {
                        throw InvalidParameter("Trying to instantiate unknown TransformationChecker " + ne.first);
                    }
                    noneLeft(p, ne.first);
                }
                if (!counterSeen) c.checker_order = 0;
            } else if (key == "inspector" || key == "logger") {
                // NullInspector / loggers: no device counterpart, nothing to configure
            } else if (key == "readingDataPointsFilters" || key == "referenceDataPointsFilters") {
                for (const auto& e : v.list) {
                    auto ne = namedEntry(e, "icp." + key);
                    if (ne.first != "IdentityDataPointsFilter")
                        throw InvalidParameter("icp." + key + ": " + ne.first + " is not implemented inside the registration (put it in the `input:` chain)");
                }
            } else {
                throw InvalidParameter("icp: unknown key " + key);
            }
        }
    }

    if (const yaml::Node* post = root.find("post")) {  // Mapper.cpp:90-98
        if (!post->isNull() && post->kind != yaml::Node::List) throw InvalidParameter("post: expected a list of DataPointsFilters");
        for (const auto& e : post->list) {
            auto ne = namedEntry(e, "post");
            Parameters& p = ne.second;
            if (ne.first == "SurfaceNormalDataPointsFilter") {
                cfg.post.surfaceNormalKnn = (int)take(p, "knn", 5.f);
                take(p, "epsilon", 0.f);
                take(p, "keepNormals", 1.f);
            } else if (ne.first == "CutAtDescriptorThresholdDataPointsFilter") {
                if (takeStr(p, "descName", "none") != "probabilityDynamic")
                    throw InvalidParameter("CutAtDescriptorThresholdDataPointsFilter: only descName probabilityDynamic is implemented on this path");
                cfg.post.cutAtThreshold = true;
                cfg.post.cutUseLargerThan = take(p, "useLargerThan", 1.f) != 0.f;
                cfg.post.cutThreshold = take(p, "threshold", 0.f);
            } else if (ne.first == "IdentityDataPointsFilter") {
            } else {
                throw InvalidParameter("Trying to instantiate unknown DataPointsFilter " + ne.first + " (post chain of the B200 path: SurfaceNormal, CutAtDescriptorThreshold)");
            }
            noneLeft(p, ne.first);
        }
    }

    if (const yaml::Node* mapper = root.find("mapper")) {  // Mapper.cpp:100-176
        if (mapper->kind != yaml::Node::Map) throw InvalidParameter("mapper: expected a map");
        for (const auto& kv : mapper->map) {
            if (kv.first == "updateCondition") {
                const yaml::Node* type = kv.second.find("type");
                const yaml::Node* value = kv.second.find("value");
                if (type) cfg.mapUpdateCondition = type->scalar;
                if (value) cfg.mapUpdateValue = num(*value, "mapper.updateCondition.value");
            } else if (kv.first == "sensorMaxRange") {
                cfg.sensorMaxRange = num(kv.second, "mapper.sensorMaxRange");
            } else if (kv.first == "mapperModule") {
                if (kv.second.kind != yaml::Node::List) throw InvalidParameter("mapper.mapperModule: expected a list");
                for (const auto& e : kv.second.list) cfg.mapperModules.push_back(namedEntry(e, "mapper.mapperModule"));
            } else {
                throw InvalidParameter("mapper: unknown key " + kv.first);
            }
        }
    }
    for (const auto& kv : root.map)
        if (kv.first != "input" && kv.first != "icp" && kv.first != "post" && kv.first != "mapper") throw InvalidParameter("unknown top-level key " + kv.first);
    return cfg;
}

inline MapperConfig loadYamlConfig(const std::string& configFilePath, bool is3D) {
    std::ifstream in(configFilePath);
    if (!in.is_open()) throw std::runtime_error("Cannot open config file: " + configFilePath);  // Mapper.cpp:61-65
    return loadYamlConfig(in, is3D);
}

}  // namespace norlab_icp_mapper_b200
