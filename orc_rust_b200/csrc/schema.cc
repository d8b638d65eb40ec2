// Reader options, column projection and the Arrow schema (src/schema.rs:503-577, src/arrow_reader.rs:70-198),
// with_schema checks (src/array_decoder/mod.rs:390-511).
#include "job_internal.h"

namespace orcb {

ReadOptions ReadOptions::from_c(const OrcbReadOptions* o) {
    ReadOptions r;
    if (!o) return r;
    r.device = o->device;
    r.batch_size = o->batch_size ? o->batch_size : 8192;
    if (o->projection_names) {
        r.project_all = false;
        for (uint32_t i = 0; i < o->n_projection; i++) r.projection.emplace_back(o->projection_names[i]);
    }
    r.range_start = o->range_start;
    r.range_end = o->range_end;
    r.timestamp_unit = o->timestamp_unit;
    r.use_row_index = !(o->flags & 1u);
    r.device_resident = o->device_resident != 0;
    r.max_stripes_per_launch = o->max_stripes_per_launch;
    r.stream = (cudaStream_t)o->cuda_stream;
    r.own_stream = o->cuda_stream == nullptr;
    r.shard_index = o->stripe_shard_index;
    r.shard_count = o->stripe_shard_count ? o->stripe_shard_count : 1;
    r.waves = o->waves;
    return r;
}

// ------------------------------------------------------------------------------------------------
// schema (src/schema.rs:503-577, src/arrow_reader.rs:182-198)
// ------------------------------------------------------------------------------------------------
// One column of the type tree as the decoder sees it (any depth; `hint`: with_schema's timestamp variant, -1 = none)
OutColumn column_info(const FileMeta& fm, uint32_t col_id, const std::string& name, const ReadOptions& opt, int hint) {
    static const char* ts_fmt[4] = {"tsn:", "tsu:", "tsm:", "tss:"};
    if (col_id >= fm.types.size()) fail(ORCB_UNEXPECTED, "Column index out of bounds");
    const OrcType& t = fm.types[col_id];
    OutColumn c;
    c.name = name;
    c.col_id = col_id;
    c.kind = t.kind;
    switch (t.kind) {
        case T_BOOLEAN: c.format = "b"; c.width = 0; break;
        case T_BYTE: c.format = "c"; c.width = 1; break;
        case T_SHORT: c.format = "s"; c.width = 2; break;
        case T_INT: c.format = "i"; c.width = 4; break;
        case T_LONG: c.format = "l"; c.width = 8; break;
        case T_FLOAT: c.format = "f"; c.width = 4; break;
        case T_DOUBLE: c.format = "g"; c.width = 8; break;
        case T_STRING: case T_VARCHAR: case T_CHAR: c.format = "u"; break;
        case T_BINARY: c.format = "z"; break;
        case T_DATE: c.format = "tdD"; c.width = 4; break;
        case T_DECIMAL:
            c.precision = t.precision;
            c.scale = t.scale;
            // arrow validates Decimal128 precision/scale (array_decoder/decimal.rs:99)
            if (t.precision == 0 || t.precision > 38 || t.scale > t.precision)
                fail(ORCB_ARROW, "invalid Decimal128 precision/scale for column " + c.name);
            c.format = "d:" + std::to_string(t.precision) + "," + std::to_string(t.scale);
            c.width = 16;
            break;
        case T_TIMESTAMP: case T_TIMESTAMP_INSTANT: {
            c.ts_unit = hint >= 0 && hint <= 3 ? hint : opt.timestamp_unit;
            c.ts_decimal = hint == 4;
            if (c.ts_decimal) {
                c.format = "d:38,9";
                c.width = 16;
                c.ts_unit = 0;
            } else {
                c.format = std::string(ts_fmt[c.ts_unit]) + (t.kind == T_TIMESTAMP_INSTANT ? "UTC" : "");
                c.width = 8;
            }
            break;
        }
        // nested types (src/schema.rs:530-577); the children are columns of their own
        case T_STRUCT:
            if (t.subtypes.size() != t.field_names.size())
                fail(ORCB_UNEXPECTED, "Struct type must have matching lengths for subtypes and field names lists");
            c.format = "+s";
            c.child_ids = t.subtypes;
            break;
        case T_LIST:
            if (t.subtypes.size() != 1) fail(ORCB_UNEXPECTED, "List type must have exactly one subtype");
            c.format = "+l";
            c.child_ids = t.subtypes;
            break;
        case T_MAP:
            if (t.subtypes.size() != 2) fail(ORCB_UNEXPECTED, "Map type must have exactly two subtypes");
            c.format = "+m";
            c.child_ids = t.subtypes;
            break;
        case T_UNION: {
            if (t.subtypes.empty() || t.subtypes.size() > 127) fail(ORCB_UNEXPECTED, "Union type must have 1..127 variants");
            c.format = "+us:";
            for (size_t i = 0; i < t.subtypes.size(); i++) c.format += (i ? "," : "") + std::to_string(i);
            c.child_ids = t.subtypes;
            break;
        }
        default: fail(ORCB_UNEXPECTED, "unknown ORC type kind " + std::to_string(t.kind));
    }
    for (uint32_t ch : c.child_ids)
        if (ch >= fm.types.size() || ch <= col_id) fail(ORCB_UNEXPECTED, "Column index out of bounds");
    return c;
}

std::vector<OutColumn> project_columns(const FileMeta& fm, const ReadOptions& opt) {
    std::vector<OutColumn> out;
    if (opt.timestamp_unit < 0 || opt.timestamp_unit > 3) fail(ORCB_INVALID_ARGUMENT, "timestamp_unit must be 0..3");
    for (auto& rc : fm.root_columns) {
        if (!opt.project_all &&
            std::find(opt.projection.begin(), opt.projection.end(), rc.first) == opt.projection.end())
            continue;
        const int hint = out.size() < opt.ts_hint.size() ? opt.ts_hint[out.size()] : -1;
        out.push_back(column_info(fm, rc.second, rc.first, opt, hint));
    }
    return out;
}

bool has_nested_columns(const std::vector<OutColumn>& cols) {
    for (auto& c : cols)
        if (!c.child_ids.empty()) return true;
    return false;
}

void apply_schema_hints(const FileMeta& fm, ReadOptions& opt, const ArrowSchema* schema) {
    if (!schema) return;
    opt.ts_hint.clear();
    const std::vector<OutColumn> cols = project_columns(fm, opt);
    if (!schema->format || strcmp(schema->format, "+s") != 0) fail(ORCB_INVALID_ARGUMENT, "with_schema: the schema must be a struct");
    if ((size_t)schema->n_children != cols.size())
        fail(ORCB_MISMATCHED_SCHEMA, "with_schema: " + std::to_string(schema->n_children) + " fields for " +
                                         std::to_string(cols.size()) + " projected columns");
    std::vector<int> hints(cols.size(), -1);
    for (size_t i = 0; i < cols.size(); i++) {
        const std::string fmt = schema->children[i]->format ? schema->children[i]->format : "";
        const OutColumn& c = cols[i];
        auto mismatch = [&]() {
            fail(ORCB_MISMATCHED_SCHEMA, "column '" + c.name + "' (ORC type kind " + std::to_string(c.kind) + ") cannot be read as Arrow '" + fmt + "'");
        };
        if (c.kind == T_TIMESTAMP || c.kind == T_TIMESTAMP_INSTANT) {
            if (fmt == "d:38,9") { hints[i] = 4; continue; }
            static const char units[4] = {'n', 'u', 'm', 's'};
            int unit = -1;
            if (fmt.size() >= 4 && fmt[0] == 't' && fmt[1] == 's' && fmt[3] == ':')
                for (int u = 0; u < 4; u++)
                    if (fmt[2] == units[u]) unit = u;
            if (unit < 0) mismatch();
            const std::string tz = fmt.substr(4);
            if (c.kind == T_TIMESTAMP) {
                if (!tz.empty()) mismatch();  // new_timestamp_decoder only takes Timestamp(_, None)
            } else {
                if (tz.empty()) mismatch();
                if (tz != "UTC") fail(ORCB_UNSUPPORTED_TYPE_VARIANT, "Non-UTC Arrow timestamps");  // timestamp.rs:214-217
            }
            hints[i] = unit;
        } else if (fmt != c.format) {
            mismatch();
        }
    }
    opt.ts_hint = hints;
}

namespace {
struct SchemaPriv {
    std::string format, name, metadata;
    std::vector<ArrowSchema> child_store;
    std::vector<ArrowSchema*> child_ptrs;
};
void release_schema(ArrowSchema* s) {
    if (!s || !s->release) return;
    for (int64_t i = 0; i < s->n_children; i++)
        if (s->children[i]->release) s->children[i]->release(s->children[i]);
    delete (SchemaPriv*)s->private_data;
    s->release = nullptr;
}
void fill_schema(ArrowSchema* s, const std::string& fmt, const std::string& name, int64_t flags) {
    auto* p = new SchemaPriv();
    p->format = fmt;
    p->name = name;
    s->format = p->format.c_str();
    s->name = p->name.c_str();
    s->metadata = nullptr;
    s->flags = flags;
    s->n_children = 0;
    s->children = nullptr;
    s->dictionary = nullptr;
    s->release = release_schema;
    s->private_data = p;
}
}  // namespace

// Exception safety: a node's children are linked (zeroed, release == nullptr) before any of them is built, so that
// releasing the root frees whatever exists when a damaged file makes column_info throw half-way.
static void link_children(ArrowSchema* s, size_t n) {
    auto* p = (SchemaPriv*)s->private_data;
    p->child_store.assign(n, ArrowSchema{});
    p->child_ptrs.resize(n);
    for (size_t i = 0; i < n; i++) p->child_ptrs[i] = &p->child_store[i];
    s->n_children = (int64_t)n;
    s->children = p->child_ptrs.data();
}

static void export_column_schema(const FileMeta& fm, const OutColumn& c, const ReadOptions& opt, const std::string& name,
                                 int64_t flags, ArrowSchema* out) {
    fill_schema(out, c.format, name, flags);
    if (c.child_ids.empty()) return;
    const OrcType& t = fm.types[c.col_id];
    const int64_t NULLABLE = 2;
    if (c.kind == T_MAP) {
        // Map(entries: Struct(keys non-null, values nullable), non-null), not sorted (src/schema.rs:548-557)
        link_children(out, 1);
        ArrowSchema* entries = out->children[0];
        fill_schema(entries, "+s", "entries", 0);
        link_children(entries, 2);
        export_column_schema(fm, column_info(fm, c.child_ids.at(0), "keys", opt, -1), opt, "keys", 0, entries->children[0]);
        export_column_schema(fm, column_info(fm, c.child_ids.at(1), "values", opt, -1), opt, "values", NULLABLE, entries->children[1]);
    } else {
        link_children(out, c.child_ids.size());
        for (size_t i = 0; i < c.child_ids.size(); i++) {
            const std::string nm = c.kind == T_STRUCT ? t.field_names.at(i) : c.kind == T_LIST ? std::string("item") : "_union_" + std::to_string(i);
            export_column_schema(fm, column_info(fm, c.child_ids[i], nm, opt, -1), opt, nm, NULLABLE, out->children[i]);
        }
    }
}

void export_schema(const FileMeta& fm, const std::vector<OutColumn>& cols, const ReadOptions& opt, ArrowSchema* out) {
    fill_schema(out, "+s", "", 0);
    try {
        auto* p = (SchemaPriv*)out->private_data;
        if (!fm.user_metadata.empty()) {
            std::string& m = p->metadata;
            auto put32 = [&](int32_t v) { m.append((const char*)&v, 4); };
            put32((int32_t)fm.user_metadata.size());
            for (auto& kv : fm.user_metadata) {
                put32((int32_t)kv.first.size());
                m.append(kv.first);
                put32((int32_t)kv.second.size());
                m.append(kv.second);
            }
            out->metadata = p->metadata.data();
        }
        link_children(out, cols.size());
        for (size_t i = 0; i < cols.size(); i++)
            export_column_schema(fm, cols[i], opt, cols[i].name, 2 /* ARROW_FLAG_NULLABLE: src/schema.rs:131 */, out->children[i]);
    } catch (...) {
        out->release(out);
        throw;
    }
}

}  // namespace orcb
