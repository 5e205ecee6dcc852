// foamfile.cpp — OpenFOAM ASCII on-disk formats (include/rheo_io.h): polyMesh directories and vol<Type>Field files,
// plain or gzip.  EXT-OF9 behaviour restated: FoamFile header, List<T> syntax "N ( ... )" and "N { v }", polyBoundaryMesh
// entries, dictionary lookup with regular-expression keywords (exact match first, then patterns, last one wins).
#include <zlib.h>

#include <algorithm>
#include <array>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <regex>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "host_mesh.hpp"
#include "rheo_io.h"

namespace {

// ------------------------------------------------------------------ file <-> string (gz transparent)
bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }

bool slurp(const std::string& path, std::string& out) {
    std::string p = path;
    if (!file_exists(p)) { if (file_exists(p + ".gz")) p += ".gz"; else return false; }
    gzFile f = gzopen(p.c_str(), "rb");   // reads plain files too
    if (!f) return false;
    out.clear();
    char buf[1 << 16];
    int n;
    while ((n = gzread(f, buf, sizeof buf)) > 0) out.append(buf, (size_t)n);
    gzclose(f);
    return n == 0;
}

bool spit(const std::string& path, const std::string& data, bool gz) {
    if (gz) {
        gzFile f = gzopen((path + ".gz").c_str(), "wb");
        if (!f) return false;
        const bool ok = gzwrite(f, data.data(), (unsigned)data.size()) == (int)data.size();
        return gzclose(f) == Z_OK && ok;
    }
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    return fclose(f) == 0 && ok;
}

// ------------------------------------------------------------------ tokenizer
struct Tok { enum Kind { Punct, Word, Str } kind; std::string s; };

bool tokenize(const std::string& in, std::vector<Tok>& out) {
    size_t i = 0, n = in.size();
    while (i < n) {
        const char c = in[i];
        if (std::isspace((unsigned char)c)) { ++i; continue; }
        if (c == '/' && i + 1 < n && in[i + 1] == '/') { while (i < n && in[i] != '\n') ++i; continue; }
        if (c == '/' && i + 1 < n && in[i + 1] == '*') {
            const size_t e = in.find("*/", i + 2);
            if (e == std::string::npos) return false;
            i = e + 2; continue;
        }
        if (c == '#' && i + 1 < n && in[i + 1] == '{') {   // verbatim code block of a coded boundary condition: #{ ... #}
            const size_t e = in.find("#}", i + 2);
            if (e == std::string::npos) return false;
            out.push_back({Tok::Str, in.substr(i + 2, e - i - 2)});
            i = e + 2; continue;
        }
        if (c == '(' || c == ')' || c == '{' || c == '}' || c == ';') { out.push_back({Tok::Punct, std::string(1, c)}); ++i; continue; }
        if (c == '"') {
            size_t e = i + 1;
            while (e < n && in[e] != '"') { if (in[e] == '\\') ++e; ++e; }
            if (e >= n) return false;
            out.push_back({Tok::Str, in.substr(i + 1, e - i - 1)});
            i = e + 1; continue;
        }
        size_t e = i;
        int depth = 0;   // words may contain balanced <> (List<symmTensor>) and [] is handled as word characters
        if (c == '$' && i + 1 < n && in[i + 1] == '{') {   // ${name}
            const size_t close = in.find('}', i + 2);
            if (close == std::string::npos) return false;
            out.push_back({Tok::Word, in.substr(i, close - i + 1)});
            i = close + 1; continue;
        }
        // a word that starts with a letter may hold balanced parentheses: div(phi,theta), ddt(theta), grad(U)
        // (EXT-OF9 ISstream::readWordToken); numbers never do: "4(0 1 2 3)" is the label 4 followed by a list
        const bool alpha = std::isalpha((unsigned char)c) || c == '_';
        int pdepth = 0;
        while (e < n) {
            const char d = in[e];
            if (d == '<') ++depth;
            if (d == '>') --depth;
            if (alpha && d == '(' && e > i) { ++pdepth; ++e; continue; }
            if (pdepth > 0 && d == ')') { --pdepth; ++e; continue; }
            if (depth == 0 && pdepth == 0 && (std::isspace((unsigned char)d) || d == '(' || d == ')' || d == '{' || d == '}' || d == ';' || d == '"')) break;
            if (pdepth > 0 && (d == ';' || d == '{' || d == '}')) break;   // unbalanced: give up on the parenthesis
            ++e;
        }
        out.push_back({Tok::Word, in.substr(i, e - i)});
        i = e;
    }
    return true;
}

// ------------------------------------------------------------------ dictionary
struct Dict;
struct Entry {
    std::string key;
    bool pattern = false;              // quoted keyword = regular expression
    std::vector<Tok> value;            // tokens up to ';' (primitive entry)
    std::shared_ptr<Dict> sub;         // or a sub-dictionary
};
struct Dict {
    std::vector<Entry> entries;
    const Entry* find_exact(const std::string& k) const {
        for (const Entry& e : entries) if (!e.pattern && e.key == k) return &e;
        return nullptr;
    }
    // EXT-OF9 dictionary::lookupEntryPtr with patternMatch: exact keyword, else patterns searched last-to-first
    const Entry* lookup(const std::string& k) const {
        if (const Entry* e = find_exact(k)) return e;
        for (auto it = entries.rbegin(); it != entries.rend(); ++it) {
            if (!it->pattern) continue;
            try { if (std::regex_match(k, std::regex(it->key, std::regex::extended))) return &*it; } catch (...) {}
        }
        return nullptr;
    }
};

bool parse_dict(const std::vector<Tok>& t, size_t& i, Dict& d, bool top) {
    while (i < t.size()) {
        if (t[i].kind == Tok::Punct && t[i].s == "}") { if (top) return false; ++i; return true; }
        if (t[i].kind == Tok::Punct) return false;
        if (t[i].kind == Tok::Word && t[i].s[0] == '#') {   // #include "file", #includeEtc "...", #inputMode merge: not followed, not needed for the values
            i += 2;
            continue;
        }
        Entry e;
        e.key = t[i].s; e.pattern = t[i].kind == Tok::Str;
        ++i;
        if (i < t.size() && t[i].kind == Tok::Punct && t[i].s == "{") {
            ++i;
            e.sub = std::make_shared<Dict>();
            if (!parse_dict(t, i, *e.sub, false)) return false;
        } else {
            int depth = 0;
            while (i < t.size()) {
                if (t[i].kind == Tok::Punct) {
                    if (t[i].s == "(") ++depth;
                    else if (t[i].s == ")") --depth;
                    else if (t[i].s == ";" && depth == 0) break;
                    else if (t[i].s == "{" || t[i].s == "}") { if (depth == 0) return false; }
                }
                e.value.push_back(t[i]);
                ++i;
            }
            if (i >= t.size()) return false;
            ++i;   // ';'
        }
        d.entries.push_back(std::move(e));
    }
    return top;
}

// $name -> the tokens of entry `name`, looked up from the innermost scope outwards (EXT-OF9 dictionary variable expansion)
std::vector<Tok> expand(const std::vector<Tok>& v, const std::vector<const Dict*>& scopes, int depth = 0) {
    std::vector<Tok> out;
    for (const Tok& t : v) {
        if (t.kind == Tok::Word && t.s.size() > 1 && t.s[0] == '$' && depth < 8) {
            std::string name = t.s.substr(1);
            if (!name.empty() && name[0] == '{' && name.back() == '}') name = name.substr(1, name.size() - 2);
            const Entry* e = nullptr;
            for (auto it = scopes.rbegin(); it != scopes.rend() && !e; ++it) e = (*it)->find_exact(name);
            if (e && !e->sub) { const std::vector<Tok> x = expand(e->value, scopes, depth + 1); out.insert(out.end(), x.begin(), x.end()); continue; }
        }
        out.push_back(t);
    }
    return out;
}

bool to_double(const std::string& s, double& v) { char* e = nullptr; v = std::strtod(s.c_str(), &e); return e && *e == 0 && !s.empty(); }
bool to_long(const std::string& s, long& v) { char* e = nullptr; v = std::strtol(s.c_str(), &e, 10); return e && *e == 0 && !s.empty(); }

// value tokens of "uniform (a b c)" / "uniform s" / "nonuniform List<T> N ( (..) (..) )" / "nonuniform List<T> N{v}" / "N ( ... )"
struct FieldValue { bool uniform = true; int ncomp = 0; std::vector<double> data; long count = 0; };

bool parse_value(const std::vector<Tok>& v, FieldValue& out) {
    size_t i = 0;
    if (v.empty()) return false;
    bool nonuni = false;
    if (v[i].s == "uniform") { ++i; }
    else if (v[i].s == "nonuniform") { nonuni = true; ++i; if (i < v.size() && v[i].kind == Tok::Word && v[i].s.rfind("List<", 0) == 0) ++i; }
    out.uniform = !nonuni;
    auto read_group = [&](std::vector<double>& dst, int& nc) -> bool {   // "(a b c)" or a bare scalar
        if (i >= v.size()) return false;
        if (v[i].kind == Tok::Punct && v[i].s == "(") {
            ++i; int c = 0;
            while (i < v.size() && !(v[i].kind == Tok::Punct && v[i].s == ")")) { double x; if (!to_double(v[i].s, x)) return false; dst.push_back(x); ++c; ++i; }
            if (i >= v.size()) return false;
            ++i; nc = c; return true;
        }
        double x; if (!to_double(v[i].s, x)) return false;
        dst.push_back(x); ++i; nc = 1; return true;
    };
    if (!nonuni) { out.count = 1; return read_group(out.data, out.ncomp) && i == v.size(); }
    // N ( ... )  |  N{v}  (the tokenizer splits "N{v}" as N { v } only through parse_dict; here it arrives as N ( ... ) or "0()" )
    long n = 0;
    if (i >= v.size()) return false;
    if (!to_long(v[i].s, n)) return false;
    ++i;
    out.count = n;
    if (i >= v.size()) return n == 0;
    if (!(v[i].kind == Tok::Punct && v[i].s == "(")) return false;
    ++i;
    out.data.reserve((size_t)n * 6);
    for (long q = 0; q < n; ++q) { int nc = 0; if (!read_group(out.data, nc)) return false; if (q == 0) out.ncomp = nc; else if (nc != out.ncomp) return false; }
    if (i >= v.size() || !(v[i].kind == Tok::Punct && v[i].s == ")")) return false;
    return true;
}

int ncomp_of_class(const std::string& cls) {
    if (cls.find("SymmTensor") != std::string::npos) return 6;
    if (cls.find("Tensor") != std::string::npos) return 9;
    if (cls.find("Vector") != std::string::npos) return 3;
    if (cls.find("Scalar") != std::string::npos) return 1;
    return 0;
}

std::string banner(const std::string& cls, const std::string& location, const std::string& object, const std::string& note = "") {
    std::string s =
        "/*--------------------------------*- C++ -*----------------------------------*\\\n"
        "  =========                 |\n"
        "  \\\\      /  F ield         | written by rheo-b200 (OpenFOAM-9 ASCII format)\n"
        "   \\\\    /   O peration     |\n"
        "    \\\\  /    A nd           |\n"
        "     \\\\/     M anipulation  |\n"
        "\\*---------------------------------------------------------------------------*/\n"
        "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       " + cls + ";\n";
    if (!note.empty()) s += "    note        \"" + note + "\";\n";
    if (!location.empty()) s += "    location    \"" + location + "\";\n";
    s += "    object      " + object + ";\n}\n// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n";
    return s;
}

void put_g17(std::string& s, double v) { char b[40]; snprintf(b, sizeof b, "%.17g", v); s += b; }

// skip the FoamFile header dictionary of a tokenised file; returns index of the first body token; fills class/object/note
bool skip_header(const std::vector<Tok>& t, size_t& i, std::string* cls, std::string* object, std::string* note) {
    i = 0;
    if (t.size() < 2 || t[0].s != "FoamFile") return false;
    i = 1;
    if (!(t[i].kind == Tok::Punct && t[i].s == "{")) return false;
    ++i;
    Dict d;
    if (!parse_dict(t, i, d, false)) return false;
    auto get = [&](const char* k, std::string* dst) { if (!dst) return; if (const Entry* e = d.find_exact(k)) if (!e->value.empty()) *dst = e->value[0].s; };
    get("class", cls); get("object", object); get("note", note);
    return true;
}

}  // namespace

struct RheoFoamDict { Dict root; };

struct RheoFoamField {
    std::string cls, object;
    int ncomp = 0;
    FieldValue internal;
    Dict body;        // the whole file (scope of $variables)
    Dict boundary;
};

namespace {

bool patch_bc_code(const std::string& type, int32_t& code) {
    if (type == "fixedValue") code = RHEO_BC_FIXED_VALUE;
    else if (type == "zeroGradient") code = RHEO_BC_ZERO_GRADIENT;
    else if (type == "linearExtrapolation") code = RHEO_BC_LINEAR_EXTRAPOLATION;
    else if (type == "empty") code = RHEO_BC_EMPTY;
    else if (type == "processor") code = RHEO_BC_PROCESSOR;
    else return false;
    return true;
}

std::string patch_name_of(const RheoHostMesh* m, int p) {
    if (p < (int)m->patch_names.size() && !m->patch_names[p].empty()) return m->patch_names[p];
    if (m->patches[p].type == RHEO_PATCH_PROCESSOR)   // EXT-OF9 processorPolyPatch::newName
        return "procBoundary" + std::to_string(std::max(m->my_rank, 0)) + "to" + std::to_string(m->patches[p].nbr_rank);
    return "patch" + std::to_string(p);
}

// points / faces of a generated tensor grid (host_mesh.hpp provenance): points are numbered on first use
bool grid_points_faces(const RheoHostMesh& m, std::vector<double>& pts, std::vector<int32_t>& fstart, std::vector<int32_t>& fpts) {
    if (!m.has_grid) return false;
    std::map<std::array<long, 3>, int32_t> ids;
    const auto find_idx = [](const std::vector<double>& v, double x) { return (long)(std::lower_bound(v.begin(), v.end(), x) - v.begin()); };
    fstart.assign(1, 0);
    for (int f = 0; f < m.n_faces; ++f) {
        double p[4][3];
        rheo::grid_face_points(m, &m.cell_ijk[3 * (size_t)m.owner[f]], m.face_dir[f], p);
        for (int q = 0; q < 4; ++q) {
            const std::array<long, 3> key{find_idx(m.xs, p[q][0]), find_idx(m.ys, p[q][1]), find_idx(m.zs, p[q][2])};
            auto it = ids.find(key);
            if (it == ids.end()) { it = ids.emplace(key, (int32_t)(pts.size() / 3)).first; pts.insert(pts.end(), {p[q][0], p[q][1], p[q][2]}); }
            fpts.push_back(it->second);
        }
        fstart.push_back((int32_t)fpts.size());
    }
    return true;
}

}  // namespace

extern "C" {

RheoHostMesh* rheo_io_read_polymesh(const char* dir) {
    if (!dir) { rheo::set_error("rheo_io_read_polymesh: null directory"); return nullptr; }
    const std::string d(dir);
    auto fail = [&](const std::string& msg) -> RheoHostMesh* { rheo::set_error("rheo_io_read_polymesh(" + d + "): " + msg); return nullptr; };
    auto load = [&](const char* name, std::vector<Tok>& t, size_t& i) -> bool {
        std::string txt;
        if (!slurp(d + "/" + name, txt)) return false;
        if (!tokenize(txt, t)) return false;
        return skip_header(t, i, nullptr, nullptr, nullptr);
    };
    auto m = std::make_unique<RheoHostMesh>();
    // ---- points
    {
        std::vector<Tok> t; size_t i;
        if (!load("points", t, i)) return fail("cannot read points");
        long n;
        if (i >= t.size() || !to_long(t[i].s, n) || t[i + 1].s != "(") return fail("points: bad list header");
        i += 2;
        m->points.resize(3 * (size_t)n);
        for (long q = 0; q < n; ++q) {
            if (i + 4 >= t.size() || t[i].s != "(") return fail("points: bad entry");
            for (int c = 0; c < 3; ++c) if (!to_double(t[i + 1 + c].s, m->points[3 * (size_t)q + c])) return fail("points: bad number");
            i += 5;
        }
    }
    // ---- faces
    {
        std::vector<Tok> t; size_t i;
        if (!load("faces", t, i)) return fail("cannot read faces");
        long n;
        if (i >= t.size() || !to_long(t[i].s, n) || t[i + 1].s != "(") return fail("faces: bad list header");
        i += 2;
        m->face_start.assign(1, 0);
        for (long q = 0; q < n; ++q) {
            long k;
            if (i + 1 >= t.size() || !to_long(t[i].s, k) || t[i + 1].s != "(") return fail("faces: bad entry");
            i += 2;
            for (long c = 0; c < k; ++c) { long v; if (i >= t.size() || !to_long(t[i].s, v)) return fail("faces: bad label"); m->face_pts.push_back((int32_t)v); ++i; }
            if (i >= t.size() || t[i].s != ")") return fail("faces: unterminated entry");
            ++i;
            m->face_start.push_back((int32_t)m->face_pts.size());
        }
        m->n_faces = (int32_t)n;
    }
    // ---- owner / neighbour
    auto read_labels = [&](const char* name, std::vector<int32_t>& dst) -> bool {
        std::vector<Tok> t; size_t i;
        if (!load(name, t, i)) return false;
        long n;
        if (i >= t.size() || !to_long(t[i].s, n) || t[i + 1].s != "(") return false;
        i += 2;
        dst.resize((size_t)n);
        for (long q = 0; q < n; ++q) { long v; if (i >= t.size() || !to_long(t[i].s, v)) return false; dst[(size_t)q] = (int32_t)v; ++i; }
        return true;
    };
    if (!read_labels("owner", m->owner)) return fail("cannot read owner");
    if (!read_labels("neighbour", m->neighbour)) return fail("cannot read neighbour");
    if ((int)m->owner.size() != m->n_faces) return fail("owner and faces differ in length");
    if (file_exists(d + "/cellProcAddressing") || file_exists(d + "/cellProcAddressing.gz")) {   // processor mesh written by decomposePar
        if (!read_labels("cellProcAddressing", m->cell_addr) || !read_labels("faceProcAddressing", m->face_addr)) return fail("cannot read the ProcAddressing files");
    }
    m->n_internal = (int32_t)m->neighbour.size();
    m->n_cells = 0;
    for (int32_t o : m->owner) m->n_cells = std::max(m->n_cells, o + 1);
    for (int32_t o : m->neighbour) m->n_cells = std::max(m->n_cells, o + 1);
    // ---- boundary
    {
        std::vector<Tok> t; size_t i;
        if (!load("boundary", t, i)) return fail("cannot read boundary");
        long n;
        if (i >= t.size() || !to_long(t[i].s, n) || t[i + 1].s != "(") return fail("boundary: bad list header");
        i += 2;
        for (long q = 0; q < n; ++q) {
            if (i + 1 >= t.size() || t[i + 1].s != "{") return fail("boundary: bad patch entry");
            const std::string name = t[i].s;
            i += 2;
            Dict pd;
            if (!parse_dict(t, i, pd, false)) return fail("boundary: bad patch dictionary " + name);
            auto word = [&](const char* k) -> std::string { const Entry* e = pd.find_exact(k); return (e && !e->value.empty()) ? e->value[0].s : std::string(); };
            RheoPatchDesc p{};
            const std::string type = word("type");
            if (type == "patch") p.type = RHEO_PATCH_PATCH;
            else if (type == "wall") p.type = RHEO_PATCH_WALL;
            else if (type == "empty") p.type = RHEO_PATCH_EMPTY;
            else if (type == "processor") p.type = RHEO_PATCH_PROCESSOR;
            else return fail("patch " + name + " has type " + type + ": only patch, wall, empty and processor patches exist on the stress-step path");
            long nf, sf;
            if (!to_long(word("nFaces"), nf) || !to_long(word("startFace"), sf)) return fail("patch " + name + ": nFaces/startFace missing");
            p.size = (int32_t)nf; p.start = (int32_t)sf; p.nbr_rank = -1;
            if (p.type == RHEO_PATCH_PROCESSOR) {
                long r;
                if (!to_long(word("neighbProcNo"), r)) return fail("processor patch " + name + ": neighbProcNo missing");
                p.nbr_rank = (int32_t)r;
                if (to_long(word("myProcNo"), r)) m->my_rank = (int32_t)r;
            }
            p.theta_bc = p.tau_bc = p.type == RHEO_PATCH_EMPTY ? RHEO_BC_EMPTY : (p.type == RHEO_PATCH_PROCESSOR ? RHEO_BC_PROCESSOR : RHEO_BC_ZERO_GRADIENT);
            m->patches.push_back(p);
            m->patch_names.push_back(name);
        }
    }
    // ---- geometry (EXT-OF9 primitiveMesh::makeFaceCentresAndAreas / makeCellCentresAndVols, surfaceInterpolation::makeWeights)
    const size_t nF = (size_t)m->n_faces;
    m->Sf.resize(3 * nF); m->Cf.resize(3 * nF);
    std::vector<double> buf;
    for (size_t f = 0; f < nF; ++f) {
        const int k = m->face_start[f + 1] - m->face_start[f];
        if (k < 3) return fail("face with fewer than 3 points");
        buf.resize(3 * (size_t)k);
        for (int q = 0; q < k; ++q) {
            const int32_t pt = m->face_pts[(size_t)m->face_start[f] + q];
            if (pt < 0 || 3 * (size_t)pt + 2 >= m->points.size()) return fail("face refers to a point that does not exist");
            for (int c = 0; c < 3; ++c) buf[3 * (size_t)q + c] = m->points[3 * (size_t)pt + c];
        }
        rheo::face_centre_area(reinterpret_cast<const double(*)[3]>(buf.data()), k, &m->Cf[3 * f], &m->Sf[3 * f]);
    }
    rheo::cell_centres_volumes(m->n_cells, m->n_faces, m->n_internal, m->owner.data(), m->neighbour.data(), m->Cf.data(), m->Sf.data(), m->C, m->V);
    rheo::linear_weights(*m);
    m->nbr_C.assign(3 * (size_t)m->n_boundary_faces(), 0.0);
    m->global_cell.resize(m->n_cells);
    for (int c = 0; c < m->n_cells; ++c) m->global_cell[c] = c;
    // ---- solved components: EXT-OF9 polyMesh::solutionD() is -1 in the directions normal to the `empty` patches, and
    // polyMesh::validComponents<symmTensor>() = solutionD x solutionD > 0, i.e. in a 2-D x-y case xz and yz drop out (zz stays:
    // (-1)(-1) = +1), which is what fvMatrix::solveSegregated loops over (the tensor-grid generator applies the same rule)
    {
        int sd[3] = {1, 1, 1};
        for (const RheoPatchDesc& p : m->patches) {
            if (p.type != RHEO_PATCH_EMPTY) continue;
            for (int f = p.start; f < p.start + p.size; ++f) {
                const double* S = &m->Sf[3 * (size_t)f];
                const double mag = std::sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
                for (int d = 0; d < 3; ++d) if (std::fabs(S[d]) > 0.999 * mag) sd[d] = -1;
            }
        }
        const int ij[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
        for (int q = 0; q < 6; ++q) m->solved[q] = sd[ij[q][0]] * sd[ij[q][1]] > 0 ? 1 : 0;
    }
    return m.release();
}

int rheo_io_write_polymesh(const RheoHostMesh* m, const char* dir, int32_t gz) {
    if (!m || !dir) { rheo::set_error("rheo_io_write_polymesh: null argument"); return 1; }
    std::vector<double> gp; std::vector<int32_t> gs, gf;
    const std::vector<double>* pts = &m->points; const std::vector<int32_t>*fs = &m->face_start, *fp = &m->face_pts;
    if (m->points.empty()) {
        if (!grid_points_faces(*m, gp, gs, gf)) { rheo::set_error("rheo_io_write_polymesh: the mesh carries no points (neither read from disk nor a generated tensor grid)"); return 1; }
        pts = &gp; fs = &gs; fp = &gf;
    }
    const std::string d(dir);
    const std::string note = "nPoints:" + std::to_string(pts->size() / 3) + "  nCells:" + std::to_string(m->n_cells) + "  nFaces:" + std::to_string(m->n_faces) +
                             "  nInternalFaces:" + std::to_string(m->n_internal);
    std::string s = banner("vectorField", "constant/polyMesh", "points");
    s += std::to_string(pts->size() / 3) + "\n(\n";
    for (size_t q = 0; q < pts->size() / 3; ++q) { s += "("; put_g17(s, (*pts)[3 * q]); s += " "; put_g17(s, (*pts)[3 * q + 1]); s += " "; put_g17(s, (*pts)[3 * q + 2]); s += ")\n"; }
    s += ")\n";
    if (!spit(d + "/points", s, gz != 0)) { rheo::set_error("rheo_io_write_polymesh: cannot write " + d + "/points"); return 1; }
    s = banner("faceList", "constant/polyMesh", "faces");
    s += std::to_string(m->n_faces) + "\n(\n";
    for (int f = 0; f < m->n_faces; ++f) {
        s += std::to_string((*fs)[f + 1] - (*fs)[f]) + "(";
        for (int q = (*fs)[f]; q < (*fs)[f + 1]; ++q) { if (q > (*fs)[f]) s += " "; s += std::to_string((*fp)[q]); }
        s += ")\n";
    }
    s += ")\n";
    if (!spit(d + "/faces", s, gz != 0)) { rheo::set_error("rheo_io_write_polymesh: cannot write faces"); return 1; }
    auto labels = [&](const char* name, const std::vector<int32_t>& v) {
        std::string o = banner("labelList", "constant/polyMesh", name, note);
        o += std::to_string(v.size()) + "\n(\n";
        for (int32_t x : v) o += std::to_string(x) + "\n";
        o += ")\n";
        return spit(d + "/" + name, o, gz != 0);
    };
    if (!labels("owner", m->owner) || !labels("neighbour", m->neighbour)) { rheo::set_error("rheo_io_write_polymesh: cannot write owner/neighbour"); return 1; }
    s = banner("polyBoundaryMesh", "constant/polyMesh", "boundary");
    s += std::to_string(m->patches.size()) + "\n(\n";
    for (size_t p = 0; p < m->patches.size(); ++p) {
        const RheoPatchDesc& pd = m->patches[p];
        const char* type = pd.type == RHEO_PATCH_WALL ? "wall" : pd.type == RHEO_PATCH_EMPTY ? "empty" : pd.type == RHEO_PATCH_PROCESSOR ? "processor" : "patch";
        s += "    " + patch_name_of(m, (int)p) + "\n    {\n        type            " + type + ";\n";
        if (pd.type == RHEO_PATCH_PROCESSOR)
            s += "        myProcNo        " + std::to_string(std::max(m->my_rank, 0)) + ";\n        neighbProcNo    " + std::to_string(pd.nbr_rank) + ";\n";
        s += "        nFaces          " + std::to_string(pd.size) + ";\n        startFace       " + std::to_string(pd.start) + ";\n    }\n";
    }
    s += ")\n";
    if (!spit(d + "/boundary", s, false)) { rheo::set_error("rheo_io_write_polymesh: cannot write boundary"); return 1; }   // the boundary file is never compressed
    if (!m->face_addr.empty() && !m->cell_addr.empty()) {   // decomposePar's addressing of a processor mesh into the undecomposed one
        std::vector<int32_t> bpa;
        int phys = 0;
        for (const RheoPatchDesc& pd : m->patches) bpa.push_back(pd.type == RHEO_PATCH_PROCESSOR ? -1 : phys++);
        if (!labels("cellProcAddressing", m->cell_addr) || !labels("faceProcAddressing", m->face_addr) || !labels("boundaryProcAddressing", bpa)) {
            rheo::set_error("rheo_io_write_polymesh: cannot write the ProcAddressing files");
            return 1;
        }
    }
    return 0;
}

int rheo_io_set_nbr_centres(RheoHostMesh* m, int32_t patch, const double* centres) {
    if (!m || !centres || patch < 0 || patch >= (int)m->patches.size() || m->patches[patch].type != RHEO_PATCH_PROCESSOR) {
        rheo::set_error("rheo_io_set_nbr_centres: not a processor patch");
        return 1;
    }
    const RheoPatchDesc& p = m->patches[patch];
    std::copy(centres, centres + 3 * (size_t)p.size, m->nbr_C.begin() + 3 * (size_t)(p.start - m->n_internal));
    // EXT-OF9 surfaceInterpolation::makeWeights on a coupled patch: w = |Sf.(Cn - Cf)| / (|Sf.(Cf - Cp)| + |Sf.(Cn - Cf)|)
    for (int32_t f = p.start; f < p.start + p.size; ++f) {
        const double* S = &m->Sf[3 * (size_t)f];
        const double* Cf = &m->Cf[3 * (size_t)f];
        const double* Cp = &m->C[3 * (size_t)m->owner[f]];
        const double* Cn = &m->nbr_C[3 * (size_t)(f - m->n_internal)];
        double so = 0, sn = 0;
        for (int c = 0; c < 3; ++c) { so += S[c] * (Cf[c] - Cp[c]); sn += S[c] * (Cn[c] - Cf[c]); }
        so = std::fabs(so); sn = std::fabs(sn);
        m->weights[f] = sn / (so + sn);
    }
    return 0;
}

int rheo_io_mesh_counts(const RheoHostMesh* m, int64_t* n_points, int64_t* n_face_points) {
    if (!m) return 1;
    if (n_points) *n_points = (int64_t)(m->points.size() / 3);
    if (n_face_points) *n_face_points = (int64_t)m->face_pts.size();
    return 0;
}

int rheo_io_patch_name(const RheoHostMesh* m, int32_t patch, char* buf, int32_t buflen) {
    if (!m || !buf || patch < 0 || patch >= (int)m->patches.size()) return 1;
    snprintf(buf, (size_t)buflen, "%s", patch_name_of(m, patch).c_str());
    return 0;
}

int rheo_io_set_patch_name(RheoHostMesh* m, int32_t patch, const char* name) {
    if (!m || !name || patch < 0 || patch >= (int)m->patches.size()) return 1;
    if (m->patch_names.size() < m->patches.size()) m->patch_names.resize(m->patches.size());
    m->patch_names[patch] = name;
    return 0;
}

RheoFoamField* rheo_io_read_field(const char* path) {
    if (!path) { rheo::set_error("rheo_io_read_field: null path"); return nullptr; }
    auto fail = [&](const std::string& msg) -> RheoFoamField* { rheo::set_error(std::string("rheo_io_read_field(") + path + "): " + msg); return nullptr; };
    std::string txt;
    if (!slurp(path, txt)) return fail("cannot read file");
    std::vector<Tok> t;
    if (!tokenize(txt, t)) return fail("unterminated comment or string");
    auto f = std::make_unique<RheoFoamField>();
    size_t i;
    if (!skip_header(t, i, &f->cls, &f->object, nullptr)) return fail("no FoamFile header");
    f->ncomp = ncomp_of_class(f->cls);
    if (!f->ncomp) return fail("class " + f->cls + " is not a vol<Type>Field this library knows");
    Dict& body = f->body;
    if (!parse_dict(t, i, body, true)) return fail("syntax error in the dictionary");
    const Entry* in = body.find_exact("internalField");
    if (!in || !parse_value(expand(in->value, {&body}), f->internal)) return fail("bad or missing internalField");
    if (f->internal.ncomp != f->ncomp && !(f->internal.count == 0)) return fail("internalField has " + std::to_string(f->internal.ncomp) + " components, class " + f->cls + " needs " + std::to_string(f->ncomp));
    const Entry* bf = body.find_exact("boundaryField");
    if (!bf || !bf->sub) return fail("no boundaryField dictionary");
    f->boundary = *bf->sub;
    return f.release();
}

void rheo_io_field_free(RheoFoamField* f) { delete f; }

RheoFoamDict* rheo_io_dict_open(const char* path) {
    if (!path) { rheo::set_error("rheo_io_dict_open: null path"); return nullptr; }
    std::string txt;
    if (!slurp(path, txt)) { rheo::set_error(std::string("rheo_io_dict_open(") + path + "): cannot read file"); return nullptr; }
    std::vector<Tok> t;
    if (!tokenize(txt, t)) { rheo::set_error(std::string("rheo_io_dict_open(") + path + "): unterminated comment or string"); return nullptr; }
    size_t i = 0;
    if (!t.empty() && t[0].s == "FoamFile") { if (!skip_header(t, i, nullptr, nullptr, nullptr)) { rheo::set_error(std::string("rheo_io_dict_open(") + path + "): bad FoamFile header"); return nullptr; } }
    auto d = std::make_unique<RheoFoamDict>();
    if (!parse_dict(t, i, d->root, true)) { rheo::set_error(std::string("rheo_io_dict_open(") + path + "): syntax error"); return nullptr; }
    return d.release();
}

void rheo_io_dict_free(RheoFoamDict* d) { delete d; }

int rheo_io_dict_lookup(const RheoFoamDict* d, const char* path, char* buf, int32_t buflen) {
    if (!d || !path || !buf || buflen < 1) { rheo::set_error("rheo_io_dict_lookup: null argument"); return 1; }
    std::vector<std::shared_ptr<Dict>> keep;          // dictionaries parsed on the way (lists of named dictionaries)
    std::vector<const Dict*> scopes{&d->root};
    const Dict* cur = &d->root;
    const Entry* e = nullptr;
    std::string p(path);
    size_t pos = 0;
    while (pos <= p.size()) {
        const size_t nx = p.find('/', pos);
        const std::string key = p.substr(pos, nx == std::string::npos ? std::string::npos : nx - pos);
        e = cur->lookup(key);
        if (!e) { rheo::set_error(std::string("rheo_io_dict_lookup: no entry ") + path); return 2; }
        if (nx == std::string::npos) break;
        if (e->sub) cur = e->sub.get();
        else {   // "( name { ... } name { ... } )": a list of named dictionaries (multiMode's `models`)
            const std::vector<Tok>& v = e->value;
            if (v.size() < 2 || v.front().s != "(" || v.back().s != ")") { rheo::set_error(std::string("rheo_io_dict_lookup: ") + key + " is not a dictionary"); return 2; }
            std::vector<Tok> inner(v.begin() + 1, v.end() - 1);
            auto sub = std::make_shared<Dict>();
            size_t j = 0;
            if (!parse_dict(inner, j, *sub, true)) { rheo::set_error(std::string("rheo_io_dict_lookup: ") + key + " is not a list of dictionaries"); return 2; }
            keep.push_back(sub);
            cur = sub.get();
        }
        scopes.push_back(cur);
        pos = nx + 1;
    }
    std::string out;
    int rc = 0;
    if (e->sub) { for (const Entry& x : e->sub->entries) { if (!out.empty()) out += ' '; out += x.key; } rc = 3; }
    else {
        const std::vector<Tok> v = expand(e->value, scopes);
        // a list of named dictionaries reports its names, like a dictionary
        bool listOfDicts = v.size() >= 2 && v.front().s == "(" && v.back().s == ")" && std::any_of(v.begin(), v.end(), [](const Tok& t) { return t.kind == Tok::Punct && t.s == "{"; });
        if (listOfDicts) {
            std::vector<Tok> inner(v.begin() + 1, v.end() - 1);
            Dict sub; size_t j = 0;
            if (parse_dict(inner, j, sub, true)) { for (const Entry& x : sub.entries) { if (!out.empty()) out += ' '; out += x.key; } rc = 3; }
        }
        if (rc == 0) for (const Tok& t : v) { if (!out.empty()) out += ' '; out += t.s; }
    }
    snprintf(buf, (size_t)buflen, "%s", out.c_str());
    return rc;
}

int rheo_io_field_info(const RheoFoamField* f, char* cls, int32_t cls_len, char* object, int32_t object_len, int32_t* n_comp, int32_t* internal_uniform,
                       int64_t* n_internal) {
    if (!f) return 1;
    if (cls) snprintf(cls, (size_t)cls_len, "%s", f->cls.c_str());
    if (object) snprintf(object, (size_t)object_len, "%s", f->object.c_str());
    if (n_comp) *n_comp = f->ncomp;
    if (internal_uniform) *internal_uniform = f->internal.uniform ? 1 : 0;
    if (n_internal) *n_internal = f->internal.uniform ? 0 : f->internal.count;
    return 0;
}

int rheo_io_field_internal(const RheoFoamField* f, int64_t n_cells, double* out) {
    if (!f || !out) { rheo::set_error("rheo_io_field_internal: null argument"); return 1; }
    if (f->internal.uniform) {
        for (int64_t c = 0; c < n_cells; ++c) for (int k = 0; k < f->ncomp; ++k) out[c * f->ncomp + k] = f->internal.data[(size_t)k];
        return 0;
    }
    if (f->internal.count != n_cells) { rheo::set_error("rheo_io_field_internal: the file holds " + std::to_string(f->internal.count) + " values, the mesh has " + std::to_string(n_cells) + " cells"); return 1; }
    std::copy(f->internal.data.begin(), f->internal.data.end(), out);
    return 0;
}

int rheo_io_field_patch(const RheoFoamField* f, const char* patch_name, int32_t n_faces, char* type, int32_t type_len, int32_t* has_value, double* values) {
    if (!f || !patch_name) { rheo::set_error("rheo_io_field_patch: null argument"); return 1; }
    const Entry* e = f->boundary.lookup(patch_name);
    if (!e || !e->sub) { rheo::set_error(std::string("rheo_io_field_patch: no boundaryField entry matches patch ") + patch_name); return 2; }
    const Entry* ty = e->sub->find_exact("type");
    if (!ty || ty->value.empty()) { rheo::set_error(std::string("rheo_io_field_patch: patch ") + patch_name + " has no type"); return 1; }
    if (type) snprintf(type, (size_t)type_len, "%s", ty->value[0].s.c_str());
    const Entry* v = e->sub->find_exact("value");
    if (has_value) *has_value = v ? 1 : 0;
    if (v && values) {
        FieldValue fv;
        if (!parse_value(expand(v->value, {&f->body, &f->boundary, e->sub.get()}), fv)) { rheo::set_error(std::string("rheo_io_field_patch: bad value entry on patch ") + patch_name); return 1; }
        if (fv.uniform) {
            if (fv.ncomp != f->ncomp) { rheo::set_error("rheo_io_field_patch: value has the wrong number of components"); return 1; }
            for (int32_t q = 0; q < n_faces; ++q) for (int k = 0; k < f->ncomp; ++k) values[(size_t)q * f->ncomp + k] = fv.data[(size_t)k];
        } else {
            if (fv.count != n_faces || (fv.count && fv.ncomp != f->ncomp)) { rheo::set_error(std::string("rheo_io_field_patch: value list of patch ") + patch_name + " does not match the patch size"); return 1; }
            std::copy(fv.data.begin(), fv.data.end(), values);
        }
    }
    return 0;
}

int rheo_io_apply_field_bcs(RheoHostMesh* m, const RheoFoamField* f, int32_t which) {
    if (!m || !f) { rheo::set_error("rheo_io_apply_field_bcs: null argument"); return 1; }
    for (size_t p = 0; p < m->patches.size(); ++p) {
        const std::string name = patch_name_of(m, (int)p);
        const Entry* e = f->boundary.lookup(name);
        if (!e || !e->sub) { rheo::set_error("rheo_io_apply_field_bcs: field " + f->object + " has no boundaryField entry for patch " + name); return 1; }
        const Entry* ty = e->sub->find_exact("type");
        int32_t code;
        if (!ty || ty->value.empty() || !patch_bc_code(ty->value[0].s, code)) {
            rheo::set_error("rheo_io_apply_field_bcs: patch " + name + " of field " + f->object + " has type " + (ty && !ty->value.empty() ? ty->value[0].s : std::string("<none>")) +
                            "; the stress step knows fixedValue, zeroGradient, linearExtrapolation, empty, processor");
            return 1;
        }
        if (code == RHEO_BC_LINEAR_EXTRAPOLATION) {
            // linearExtrapolationFvPatchField.C:72,118: `useRegression true` selects the least-squares branch (:152-219) instead of
            // the gradient branch (:101-151); only tau patches are extrapolated
            const Entry* ur = e->sub->find_exact("useRegression");
            if (ur && !ur->value.empty()) {
                const std::string& v = ur->value[0].s;
                if (v == "true" || v == "on" || v == "yes" || v == "1" || v == "y" || v == "t") code = RHEO_BC_LINEAR_EXTRAPOLATION_REG;
            }
        }
        (which == 0 ? m->patches[p].theta_bc : m->patches[p].tau_bc) = code;
    }
    return 0;
}

int rheo_io_write_field(const char* path, const char* cls, const char* object, const char* dimensions, int32_t n_comp, int64_t n_cells,
                        const double* internal, int32_t n_patches, const char* const* patch_names, const char* const* patch_types,
                        const int32_t* patch_sizes, const double* const* patch_values, int32_t gz) {
    if (!path || !cls || !object || !internal || n_comp < 1) { rheo::set_error("rheo_io_write_field: null/invalid argument"); return 1; }
    if (ncomp_of_class(cls) != n_comp) { rheo::set_error(std::string("rheo_io_write_field: class ") + cls + " does not hold " + std::to_string(n_comp) + " components"); return 1; }
    const char* tname = n_comp == 1 ? "scalar" : n_comp == 3 ? "vector" : n_comp == 6 ? "symmTensor" : "tensor";
    auto put_list = [&](std::string& s, const double* v, int64_t n) {
        s += std::string("nonuniform List<") + tname + "> " + std::to_string(n) + "\n(\n";
        for (int64_t q = 0; q < n; ++q) {
            if (n_comp > 1) s += "(";
            for (int k = 0; k < n_comp; ++k) { if (k) s += " "; put_g17(s, v[q * n_comp + k]); }
            s += n_comp > 1 ? ")\n" : "\n";
        }
        s += ")";
    };
    std::string s = banner(cls, "", object);
    s += std::string("dimensions      ") + (dimensions ? dimensions : "[0 0 0 0 0 0 0]") + ";\n\ninternalField   ";
    put_list(s, internal, n_cells);
    s += ";\n\nboundaryField\n{\n";
    for (int32_t p = 0; p < n_patches; ++p) {
        s += std::string("    ") + patch_names[p] + "\n    {\n        type            " + patch_types[p] + ";\n";
        if (patch_values && patch_values[p]) { s += "        value           "; put_list(s, patch_values[p], patch_sizes[p]); s += ";\n"; }
        s += "    }\n";
    }
    s += "}\n\n// ************************************************************************* //\n";
    if (!spit(path, s, gz != 0)) { rheo::set_error(std::string("rheo_io_write_field: cannot write ") + path); return 1; }
    return 0;
}

}  // extern "C"
