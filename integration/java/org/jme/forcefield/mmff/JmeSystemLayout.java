/*
 * JmeSystemLayout - the marshaller of the GPU drop-in: flattens what MMFF94.assignParameters leaves on an
 * IAtomContainer (MMFF94.java:111-219) into the jme_system_t of include/jme_b200.h, for Panama FFM (JDK 22+).
 *
 * Lives in package org.jme.forcefield.mmff because the property keys (MMFF94.java:59-70) and the parameter classes
 * (MMFF94Parameters.java:361-379, 432-444) are package-private; the reference jar has no module-info and is not
 * sealed, so this is legal and needs no edit of the reference's sources.
 *
 * NOT compiled in the repository's image (no JDK there).  What IS checked there: the OFF_* / SIZEOF constants below
 * against offsetof / sizeof of the C struct (tests/test_java_layout.py compiles a probe with gcc), so a drift between
 * this file and include/jme_b200.h fails the CPU test suite.
 */
package org.jme.forcefield.mmff;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_FLOAT;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Map;
import java.util.SortedMap;
import java.util.TreeMap;
import org.openscience.cdk.interfaces.IAtom;
import org.openscience.cdk.interfaces.IAtomContainer;
import org.openscience.cdk.interfaces.IBond;

final class JmeSystemLayout {

    private JmeSystemLayout() {
    }

    /* ---- byte offsets of jme_system_t (LP64), one constant per field, in declaration order ---- */
    static final long OFF_n_atoms = 0;
    static final long OFF_atom_type = 8;
    static final long OFF_charge = 16;
    static final long OFF_linear = 24;
    static final long OFF_n_vdw_types = 32;
    static final long OFF_vdw_type = 40;
    static final long OFF_vdw_alpha = 48;
    static final long OFF_vdw_N = 56;
    static final long OFF_vdw_A = 64;
    static final long OFF_vdw_G = 72;
    static final long OFF_vdw_DA = 80;
    static final long OFF_n_bonds = 88;
    static final long OFF_bond_i = 96;
    static final long OFF_bond_j = 104;
    static final long OFF_bond_kb = 112;
    static final long OFF_bond_r0 = 120;
    static final long OFF_n_angles = 128;
    static final long OFF_angle_i = 136;
    static final long OFF_angle_j = 144;
    static final long OFF_angle_k = 152;
    static final long OFF_angle_ka = 160;
    static final long OFF_angle_theta0 = 168;
    static final long OFF_stbn_kba_ijk = 176;
    static final long OFF_stbn_kba_kji = 184;
    static final long OFF_n_oops = 192;
    static final long OFF_oop_i = 200;
    static final long OFF_oop_j = 208;
    static final long OFF_oop_k = 216;
    static final long OFF_oop_l = 224;
    static final long OFF_oop_koop = 232;
    static final long OFF_n_torsions = 240;
    static final long OFF_tor_i = 248;
    static final long OFF_tor_j = 256;
    static final long OFF_tor_k = 264;
    static final long OFF_tor_l = 272;
    static final long OFF_tor_v1 = 280;
    static final long OFF_tor_v2 = 288;
    static final long OFF_tor_v3 = 296;
    static final long SIZEOF_jme_system_t = 304;

    /** The same thing as a StructLayout (what a jextract run over include/jme_b200.h produces). */
    static final StructLayout LAYOUT = MemoryLayout.structLayout(
            JAVA_LONG.withName("n_atoms"), ADDRESS.withName("atom_type"), ADDRESS.withName("charge"), ADDRESS.withName("linear"),
            JAVA_INT.withName("n_vdw_types"), MemoryLayout.paddingLayout(4),
            ADDRESS.withName("vdw_type"), ADDRESS.withName("vdw_alpha"), ADDRESS.withName("vdw_N"), ADDRESS.withName("vdw_A"),
            ADDRESS.withName("vdw_G"), ADDRESS.withName("vdw_DA"),
            JAVA_LONG.withName("n_bonds"), ADDRESS.withName("bond_i"), ADDRESS.withName("bond_j"), ADDRESS.withName("bond_kb"),
            ADDRESS.withName("bond_r0"),
            JAVA_LONG.withName("n_angles"), ADDRESS.withName("angle_i"), ADDRESS.withName("angle_j"), ADDRESS.withName("angle_k"),
            ADDRESS.withName("angle_ka"), ADDRESS.withName("angle_theta0"), ADDRESS.withName("stbn_kba_ijk"),
            ADDRESS.withName("stbn_kba_kji"),
            JAVA_LONG.withName("n_oops"), ADDRESS.withName("oop_i"), ADDRESS.withName("oop_j"), ADDRESS.withName("oop_k"),
            ADDRESS.withName("oop_l"), ADDRESS.withName("oop_koop"),
            JAVA_LONG.withName("n_torsions"), ADDRESS.withName("tor_i"), ADDRESS.withName("tor_j"), ADDRESS.withName("tor_k"),
            ADDRESS.withName("tor_l"), ADDRESS.withName("tor_v1"), ADDRESS.withName("tor_v2"), ADDRESS.withName("tor_v3"))
            .withName("jme_system_t");

    /* ---- growable primitive columns ---- */
    private static final class IntCol {
        int[] v = new int[64];
        int n;
        void add(int x) { if (n == v.length) v = java.util.Arrays.copyOf(v, 2 * n); v[n++] = x; }
        MemorySegment to(Arena a) { MemorySegment s = a.allocate(JAVA_INT, Math.max(1, n)); MemorySegment.copy(v, 0, s, JAVA_INT, 0, n); return s; }
    }

    private static final class FloatCol {
        float[] v = new float[64];
        int n;
        void add(float x) { if (n == v.length) v = java.util.Arrays.copyOf(v, 2 * n); v[n++] = x; }
        MemorySegment to(Arena a) { MemorySegment s = a.allocate(JAVA_FLOAT, Math.max(1, n)); MemorySegment.copy(v, 0, s, JAVA_FLOAT, 0, n); return s; }
    }

    /**
     * Build a jme_system_t in {@code arena} from a container on which MMFF94.assignParameters has run.  Every tabulated
     * parameter stays a binary32 (the reference holds them as float / Float and widens at use; the C side widens the
     * same way), atom indices are IAtom.getIndex().
     */
    @SuppressWarnings("unchecked")
    static MemorySegment write(Arena arena, IAtomContainer mol) {
        final int n = mol.getAtomCount();
        int[] type = new int[n];
        double[] q = new double[n];
        byte[] linear = new byte[n];
        SortedMap<Integer, MMFF94Parameters.VdwParameters> rows = new TreeMap<>();
        for (IAtom a : mol.atoms()) {
            int i = a.getIndex();
            type[i] = a.getProperty(MMFF94.MMFF94_TYPE);                                        // MMFF94.java:117
            q[i] = a.getCharge();                                                               // MMFF94.java:616
            MMFF94Parameters.GeometricParameters gp = a.getProperty(MMFF94.MMFF94_PARAMETER_GEOMETRIC_PROPERTIES);   // :120
            linear[i] = (byte) (gp.ideallyLinear ? 1 : 0);
            rows.put(type[i], (MMFF94Parameters.VdwParameters) a.getProperty(MMFF94.MMFF94_VDW_INTERACTION));       // :160
        }
        // vdW rows of the types that occur (MMFF94Parameters.java:361-379)
        IntCol vType = new IntCol();
        FloatCol vAlpha = new FloatCol(), vN = new FloatCol(), vA = new FloatCol(), vG = new FloatCol();
        byte[] vDA = new byte[Math.max(1, rows.size())];
        for (Map.Entry<Integer, MMFF94Parameters.VdwParameters> e : rows.entrySet()) {
            MMFF94Parameters.VdwParameters p = e.getValue();
            vDA[vType.n] = p.DA;
            vType.add(e.getKey());
            vAlpha.add(p.alpha); vN.add(p.N); vA.add(p.A); vG.add(p.G);
        }
        // bonds: container.bonds() with their StretchParameters (MMFF94.java:222-256, MMFF94Parameters.java:432-444)
        IntCol bI = new IntCol(), bJ = new IntCol();
        FloatCol bKb = new FloatCol(), bR0 = new FloatCol();
        for (IBond b : mol.bonds()) {
            MMFF94Parameters.StretchParameters sp = b.getProperty(MMFF94.MMFF94_PARAMETER_STRETCH);
            bI.add(b.getBegin().getIndex()); bJ.add(b.getEnd().getIndex());
            bKb.add(sp.kb); bR0.add(sp.r0);
        }
        // angles + stretch-bends: the two maps share their keys (MMFF94.java:180-183); rows are Float[6]
        //   angle row  {angleType, ti, tj, tk, ka, theta0}      stretch-bend row {sbType, ti, tj, tk, kbaIJK, kbaKJI}
        HashMap<List<IAtom>, Float[]> angleMap = mol.getProperty(MMFF94.MMFF94_ANGLE_BENDING);
        HashMap<List<IAtom>, Float[]> stbnMap = mol.getProperty(MMFF94.MMFF94_STRETCH_BEND);
        IntCol aI = new IntCol(), aJ = new IntCol(), aK = new IntCol();
        FloatCol aKa = new FloatCol(), aTh = new FloatCol(), sIJK = new FloatCol(), sKJI = new FloatCol();
        for (Map.Entry<List<IAtom>, Float[]> e : angleMap.entrySet()) {
            List<IAtom> k = e.getKey();
            Float[] ang = e.getValue(), sb = stbnMap.get(k);
            aI.add(k.get(0).getIndex()); aJ.add(k.get(1).getIndex()); aK.add(k.get(2).getIndex());
            aKa.add(ang[4]); aTh.add(ang[5]);
            sIJK.add(sb[4]); sKJI.add(sb[5]);
        }
        // out-of-plane: rows {., ., ., ., koop}; null rows contribute nothing (MMFF94OutOfPlaneComponent.java:124-126)
        HashMap<List<IAtom>, Float[]> oopMap = mol.getProperty(MMFF94.MMFF94_OUT_OF_PLANE);
        IntCol oI = new IntCol(), oJ = new IntCol(), oK = new IntCol(), oL = new IntCol();
        FloatCol oKoop = new FloatCol();
        for (Map.Entry<List<IAtom>, Float[]> e : oopMap.entrySet()) {
            if (e.getValue() == null) continue;
            List<IAtom> k = e.getKey();
            oI.add(k.get(0).getIndex()); oJ.add(k.get(1).getIndex()); oK.add(k.get(2).getIndex()); oL.add(k.get(3).getIndex());
            oKoop.add(e.getValue()[4]);
        }
        // torsions: rows {tt, ti, tj, tk, tl, V1, V2, V3} (MMFF94.java:196-210); the i<->l / j<->k swap of
        // MMFF94TorsionComponent.java:117-124 is done inside libjme_b200 (it is energy-neutral)
        HashMap<List<IAtom>, Float[]> torMap = mol.getProperty(MMFF94.MMFF94_TORSION);
        IntCol tI = new IntCol(), tJ = new IntCol(), tK = new IntCol(), tL = new IntCol();
        FloatCol t1 = new FloatCol(), t2 = new FloatCol(), t3 = new FloatCol();
        for (Map.Entry<List<IAtom>, Float[]> e : torMap.entrySet()) {
            List<IAtom> k = e.getKey();
            Float[] v = e.getValue();
            tI.add(k.get(0).getIndex()); tJ.add(k.get(1).getIndex()); tK.add(k.get(2).getIndex()); tL.add(k.get(3).getIndex());
            t1.add(v[5]); t2.add(v[6]); t3.add(v[7]);
        }

        MemorySegment sys = arena.allocate(LAYOUT);
        if (sys.byteSize() != SIZEOF_jme_system_t) throw new IllegalStateException("jme_system_t layout drifted");
        MemorySegment typeSeg = arena.allocate(JAVA_INT, Math.max(1, n));
        MemorySegment.copy(type, 0, typeSeg, JAVA_INT, 0, n);
        MemorySegment qSeg = arena.allocate(JAVA_DOUBLE, Math.max(1, n));
        MemorySegment.copy(q, 0, qSeg, JAVA_DOUBLE, 0, n);
        MemorySegment linSeg = arena.allocate(JAVA_BYTE, Math.max(1, n));
        MemorySegment.copy(linear, 0, linSeg, JAVA_BYTE, 0, n);
        MemorySegment daSeg = arena.allocate(JAVA_BYTE, vDA.length);
        MemorySegment.copy(vDA, 0, daSeg, JAVA_BYTE, 0, vDA.length);

        sys.set(JAVA_LONG, OFF_n_atoms, (long) n);
        sys.set(ADDRESS, OFF_atom_type, typeSeg);
        sys.set(ADDRESS, OFF_charge, qSeg);
        sys.set(ADDRESS, OFF_linear, linSeg);
        sys.set(JAVA_INT, OFF_n_vdw_types, vType.n);
        sys.set(ADDRESS, OFF_vdw_type, vType.to(arena));
        sys.set(ADDRESS, OFF_vdw_alpha, vAlpha.to(arena));
        sys.set(ADDRESS, OFF_vdw_N, vN.to(arena));
        sys.set(ADDRESS, OFF_vdw_A, vA.to(arena));
        sys.set(ADDRESS, OFF_vdw_G, vG.to(arena));
        sys.set(ADDRESS, OFF_vdw_DA, daSeg);
        sys.set(JAVA_LONG, OFF_n_bonds, (long) bI.n);
        sys.set(ADDRESS, OFF_bond_i, bI.to(arena));
        sys.set(ADDRESS, OFF_bond_j, bJ.to(arena));
        sys.set(ADDRESS, OFF_bond_kb, bKb.to(arena));
        sys.set(ADDRESS, OFF_bond_r0, bR0.to(arena));
        sys.set(JAVA_LONG, OFF_n_angles, (long) aI.n);
        sys.set(ADDRESS, OFF_angle_i, aI.to(arena));
        sys.set(ADDRESS, OFF_angle_j, aJ.to(arena));
        sys.set(ADDRESS, OFF_angle_k, aK.to(arena));
        sys.set(ADDRESS, OFF_angle_ka, aKa.to(arena));
        sys.set(ADDRESS, OFF_angle_theta0, aTh.to(arena));
        sys.set(ADDRESS, OFF_stbn_kba_ijk, sIJK.to(arena));
        sys.set(ADDRESS, OFF_stbn_kba_kji, sKJI.to(arena));
        sys.set(JAVA_LONG, OFF_n_oops, (long) oI.n);
        sys.set(ADDRESS, OFF_oop_i, oI.to(arena));
        sys.set(ADDRESS, OFF_oop_j, oJ.to(arena));
        sys.set(ADDRESS, OFF_oop_k, oK.to(arena));
        sys.set(ADDRESS, OFF_oop_l, oL.to(arena));
        sys.set(ADDRESS, OFF_oop_koop, oKoop.to(arena));
        sys.set(JAVA_LONG, OFF_n_torsions, (long) tI.n);
        sys.set(ADDRESS, OFF_tor_i, tI.to(arena));
        sys.set(ADDRESS, OFF_tor_j, tJ.to(arena));
        sys.set(ADDRESS, OFF_tor_k, tK.to(arena));
        sys.set(ADDRESS, OFF_tor_l, tL.to(arena));
        sys.set(ADDRESS, OFF_tor_v1, t1.to(arena));
        sys.set(ADDRESS, OFF_tor_v2, t2.to(arena));
        sys.set(ADDRESS, OFF_tor_v3, t3.to(arena));
        return sys;
    }

    /** Scratch list kept so that callers may reuse the type (documentation of the row shapes above). */
    static List<String> fieldNames() {
        List<String> names = new ArrayList<>();
        LAYOUT.memberLayouts().forEach(m -> m.name().ifPresent(names::add));
        return names;
    }
}
