import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import orc_rust_b200 as ob
p = sys.argv[1] if len(sys.argv) > 1 else "tests/golden/ref_basic/nested_array.orc"
r = ob.ArrowReaderBuilder.try_new(p).build()
for b in r:
    print(b.num_rows, b.to_pydict())
