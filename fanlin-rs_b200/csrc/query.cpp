// Host-side mirror of the reference's request model, src/query.rs:3-94: the
// fields, defaults and accessor semantics the pixel-transform stage reads.  The
// Rust server keeps its own query::Query; this mirror exists so
// the C++/Python harness drives the stage with the same parameter semantics.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "plan.h"

using namespace fanlin;

namespace {

constexpr uint8_t DEFAULT_COLOR = 32;  // src/query.rs:17
constexpr uint32_t DEFAULT_QUALITY = 75;
constexpr uint32_t W_MIN = 20, W_MAX = 2000, H_MIN = 20, H_MAX = 1000;  // src/query.rs:20-21

std::string pct_decode(const std::string &s) {
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] == '+') o.push_back(' ');
        else if (s[i] == '%' && i + 2 < s.size() && isxdigit((unsigned char)s[i + 1]) && isxdigit((unsigned char)s[i + 2])) {
            o.push_back(char(strtol(s.substr(i + 1, 2).c_str(), nullptr, 16)));
            i += 2;
        } else o.push_back(s[i]);
    }
    return o;
}

// Rust's <uN as FromStr>: optional '+', then digits only, no overflow.
bool parse_uint(const std::string &s, uint64_t max, uint32_t *out) {
    size_t i = 0;
    if (!s.empty() && s[0] == '+') i = 1;
    if (i >= s.size()) return false;
    uint64_t v = 0;
    for (; i < s.size(); i++) {
        if (s[i] < '0' || s[i] > '9') return false;
        v = v * 10 + uint64_t(s[i] - '0');
        if (v > max) return false;
    }
    *out = uint32_t(v);
    return true;
}

bool parse_bool(const std::string &s, uint32_t *out) {
    if (s == "true") { *out = 1; return true; }
    if (s == "false") { *out = 0; return true; }
    return false;
}

}  // namespace

extern "C" int fanlin_query_parse(const char *query_string, fanlin_query *q) {
    if (!query_string || !q) { set_error("fanlin: null argument"); return FANLIN_EINVAL; }
    std::memset(q, 0, sizeof(*q));
    std::string s(query_string);
    const size_t qm = s.find('?');
    if (qm != std::string::npos) s = s.substr(qm + 1);
    else if (s.find("://") != std::string::npos) s.clear();  // a URL without a query
    const size_t hash = s.find('#');
    if (hash != std::string::npos) s = s.substr(0, hash);
    std::vector<std::string> pairs;
    for (size_t pos = 0; pos <= s.size();) {
        size_t amp = s.find('&', pos);
        if (amp == std::string::npos) amp = s.size();
        if (amp > pos) pairs.push_back(s.substr(pos, amp - pos));
        pos = amp + 1;
    }
    for (const std::string &pair : pairs) {
        const size_t eq = pair.find('=');
        const std::string k = pct_decode(pair.substr(0, eq));
        const std::string v = eq == std::string::npos ? std::string() : pct_decode(pair.substr(eq + 1));
        auto bad = [&](const char *what) {
            set_error(std::string("fanlin: failed to deserialize query string: ") + what + " `" + k + "=" + v + "`");
            return FANLIN_EINVAL;
        };
        auto dup = [&](uint32_t has) { return has != 0; };
        if (k == "w") { if (dup(q->has_w)) return bad("duplicate field"); if (!parse_uint(v, UINT32_MAX, &q->w)) return bad("invalid u32"); q->has_w = 1; }
        else if (k == "h") { if (dup(q->has_h)) return bad("duplicate field"); if (!parse_uint(v, UINT32_MAX, &q->h)) return bad("invalid u32"); q->has_h = 1; }
        else if (k == "rgb") {
            if (dup(q->has_rgb)) return bad("duplicate field");
            if (v.size() >= sizeof(q->rgb)) return bad("rgb too long");
            std::memcpy(q->rgb, v.c_str(), v.size() + 1);
            q->has_rgb = 1;
        }
        else if (k == "quality") { if (dup(q->has_quality)) return bad("duplicate field"); if (!parse_uint(v, 255, &q->quality)) return bad("invalid u8"); q->has_quality = 1; }
        else if (k == "blur") { if (dup(q->has_blur)) return bad("duplicate field"); if (!parse_uint(v, 255, &q->blur)) return bad("invalid u8"); q->has_blur = 1; }
        else if (k == "crop") { if (dup(q->has_crop)) return bad("duplicate field"); if (!parse_bool(v, &q->crop)) return bad("invalid bool"); q->has_crop = 1; }
        else if (k == "grayscale") { if (dup(q->has_grayscale)) return bad("duplicate field"); if (!parse_bool(v, &q->grayscale)) return bad("invalid bool"); q->has_grayscale = 1; }
        else if (k == "inverse") { if (dup(q->has_inverse)) return bad("duplicate field"); if (!parse_bool(v, &q->inverse)) return bad("invalid bool"); q->has_inverse = 1; }
        else if (k == "avif") { if (dup(q->has_avif)) return bad("duplicate field"); if (!parse_bool(v, &q->avif)) return bad("invalid bool"); q->has_avif = 1; }
        else if (k == "webp") { if (dup(q->has_webp)) return bad("duplicate field"); if (!parse_bool(v, &q->webp)) return bad("invalid bool"); q->has_webp = 1; }
        // unknown keys are ignored (src/query.rs:136-143)
    }
    return FANLIN_OK;
}

// Query::dimensions, src/query.rs:28-33: both w and h or nothing.
extern "C" int fanlin_query_dimensions(const fanlin_query *q, uint32_t *w, uint32_t *h) {
    if (!q || !q->has_w || !q->has_h) return 0;
    if (w) *w = q->w;
    if (h) *h = q->h;
    return 1;
}

// Query::fill_color, src/query.rs:35-49.
extern "C" void fanlin_query_fill_color(const fanlin_query *q, uint8_t rgb[3]) {
    rgb[0] = rgb[1] = rgb[2] = DEFAULT_COLOR;
    if (!q || !q->has_rgb) return;
    std::vector<uint8_t> parts;
    const std::string t(q->rgb);
    size_t pos = 0;
    while (parts.size() < 3) {
        size_t comma = t.find(',', pos);
        const std::string e = t.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
        uint32_t v;
        parts.push_back(parse_uint(e, 255, &v) ? uint8_t(v) : DEFAULT_COLOR);
        if (comma == std::string::npos) break;
        pos = comma + 1;
    }
    if (parts.size() != 3) return;
    rgb[0] = parts[0]; rgb[1] = parts[1]; rgb[2] = parts[2];
}

// Query::blur, src/query.rs:59-62: absent -> 0.0, else clamp(v, 10, 20).
extern "C" float fanlin_query_blur(const fanlin_query *q) {
    if (!q || !q->has_blur) return 0.0f;
    const float v = float(q->blur);
    return v < 10.0f ? 10.0f : (v > 20.0f ? 20.0f : v);
}

// Query::as_is, src/query.rs:80-87.
extern "C" int fanlin_query_as_is(const fanlin_query *q) {
    if (!q) return 1;
    return !(q->has_w && q->has_h) && fanlin_query_blur(q) == 0.0f && !(q->has_grayscale && q->grayscale) &&
           !(q->has_inverse && q->inverse) && !(q->has_avif && q->avif) && !(q->has_webp && q->webp);
}

// Query::unsupported_scale_size, src/query.rs:89-93.
extern "C" int fanlin_query_unsupported_scale_size(const fanlin_query *q) {
    const uint32_t w = q && q->has_w ? q->w : 100, h = q && q->has_h ? q->h : 100;
    return !(w >= W_MIN && w <= W_MAX) || !(h >= H_MIN && h <= H_MAX);
}

extern "C" void fanlin_job_from_query(const fanlin_query *q, int gif, fanlin_job *job) {
    if (!q || !job) return;
    uint32_t flags = 0;
    if (q->has_grayscale && q->grayscale) flags |= FANLIN_GRAYSCALE;
    if (q->has_inverse && q->inverse) flags |= FANLIN_INVERSE;
    if (q->has_crop && q->crop) flags |= FANLIN_CROP;
    job->req_w = job->req_h = 0;
    if (fanlin_query_dimensions(q, &job->req_w, &job->req_h)) flags |= FANLIN_HAS_DIMS;
    if (gif) flags |= FANLIN_TO_RGBA8;  // src/handler.rs:355
    job->flags = flags;
    job->filter = gif ? FANLIN_FILTER_NEAREST : FANLIN_FILTER_LANCZOS3;  // handler.rs:338,340 vs :233,235
    fanlin_query_fill_color(q, job->fill_rgb);
    job->blur_sigma = gif ? 0.0f : fanlin_query_blur(q);  // the GIF path never blurs
    (void)DEFAULT_QUALITY;
}
