/*
 * GpuMMFF94Component - the drop-in: one EnergyComponent that returns the whole MMFF94 energy from libjme_b200.so,
 * installed in place of the seven CPU components of MMFF94.java:98 through the reference's own registry
 * (ForceField.addEnergyComponent / removeEnergyComponent, ForceField.java:85-109).  Panama FFM, JDK 22+.
 * Not compiled in the repository's image (no JDK there); the struct marshalling is in JmeSystemLayout.java.
 */
package org.jme.forcefield.mmff;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.util.ArrayList;
import java.util.Objects;
import org.jme.forcefield.EnergyComponent;
import org.openscience.cdk.interfaces.IAtom;
import org.openscience.cdk.interfaces.IAtomContainer;

public final class GpuMMFF94Component extends EnergyComponent implements AutoCloseable {

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.libraryLookup(System.getProperty("jme.b200.lib", "libjme_b200.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }

    private static final MethodHandle CREATE = h("jme_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle DESTROY = h("jme_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle OPTIONS = h("jme_set_options", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_DOUBLE, JAVA_DOUBLE, JAVA_INT));
    private static final MethodHandle COORDS = h("jme_set_coords", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle COORD1 = h("jme_set_coord1", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_DOUBLE));
    private static final MethodHandle ENERGY = h("jme_energy", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle GRADIENT = h("jme_energy_gradient", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle DELTA = h("jme_energy_delta", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, JAVA_INT, JAVA_DOUBLE, ADDRESS));
    private static final MethodHandle INCREMENTAL = h("jme_set_incremental", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle GD = h("jme_gd_minimize", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE,
            ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle ADAM = h("jme_adam_minimize", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_DOUBLE, JAVA_DOUBLE,
            JAVA_DOUBLE, JAVA_DOUBLE, JAVA_DOUBLE, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle LASTERR = h("jme_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle CREATEERR = h("jme_last_create_error", FunctionDescriptor.of(ADDRESS));

    private Arena arena = Arena.ofShared();
    private MemorySegment handle = MemorySegment.NULL;
    private IAtomContainer bound;          // container the handle was created for
    private Object boundPairMap;           // identity of "mmff.nonBondedInteraction": a new map on every assignParameters
    private double[] shadow;               // coordinates last uploaded (to find the single moved coordinate)
    private MemorySegment xyz, total, comps, delta;
    private boolean incremental;           // opt-in: see setIncremental
    private double cachedTotal;            // energy of the coordinates in `shadow` (valid while cachedOptions match)
    private Object cachedOptions;
    private int sinceFull;

    /** Opt in to incremental energies: when one coordinate changed since the last calculateEnergy - or two: the undo of a
     *  rejected trial plus the next trial, which is what GradientDescentMinimizer.java:83-93 / AdamMinimizer.java:95-108
     *  leave between two calls - the answer is the previous energy plus jme_energy_delta (the moved atom's terms only,
     *  O(N)); every 3 N such answers, and whenever anything else changed, a full evaluation resynchronises.  The
     *  unmodified minimizers get the speed-up through the unmodified calculateEnergy. */
    public void setIncremental(boolean on) { incremental = on; }

    /** Install on an MMFF94 instance: the seven CPU components out, this one in. */
    public static GpuMMFF94Component install(MMFF94 ff) {
        for (EnergyComponent c : new ArrayList<>(ff.getEnergyComponents())) ff.removeEnergyComponent(c);
        GpuMMFF94Component gpu = new GpuMMFF94Component();
        ff.addEnergyComponent(gpu);          // calls setForceField(ff) for us (ForceField.java:96)
        return gpu;
    }

    @Override
    public double calculateEnergy(IAtomContainer mol) {
        Objects.requireNonNull(forceField);
        if (mol.isEmpty()) return 0;
        MMFF94.checkParametersAssigned(mol.getAtom(0));                       // same exception as MMFF94.java:371-375
        Object pairMap = mol.getProperty(MMFF94.MMFF94_NON_BONDED_INTERACTION);
        if (mol != bound || pairMap != boundPairMap) marshal(mol, pairMap);   // (re)assignParameters happened
        check((int) invoke(OPTIONS, handle, forceField.getCutoffDistance(), forceField.getDielectricConstant(),
                forceField.getEnergyUnit().ordinal()));                       // KCAL, KJ, JOULE = 0, 1, 2
        Object options = java.util.List.of(forceField.getCutoffDistance(), forceField.getDielectricConstant(), forceField.getEnergyUnit());
        if (incremental && shadow != null && options.equals(cachedOptions) && sinceFull < 3 * mol.getAtomCount()) {
            int n = mol.getAtomCount(), changed = 0;
            int[] which = new int[2];
            double[] now = coordinatesOf(mol);
            for (int k = 0; k < 3 * n && changed <= 2; k++) if (now[k] != shadow[k]) { if (changed < 2) which[changed] = k; changed++; }
            if (changed == 0) return cachedTotal;
            if (changed <= 2) {
                for (int c = 0; c < changed; c++) {
                    int k = which[c];
                    check((int) invoke(DELTA, handle, (long) (k / 3), k % 3, now[k], delta));
                    cachedTotal += delta.get(JAVA_DOUBLE, 0);
                    shadow[k] = now[k];
                    sinceFull++;
                }
                return cachedTotal;
            }
        }
        uploadCoordinates(mol);
        check((int) invoke(ENERGY, handle, total, comps));
        cachedTotal = total.get(JAVA_DOUBLE, 0);
        cachedOptions = options;
        sinceFull = 0;
        return cachedTotal;
    }

    private static double[] coordinatesOf(IAtomContainer mol) {
        double[] now = new double[3 * mol.getAtomCount()];
        for (IAtom a : mol.atoms()) {
            javax.vecmath.Point3d p = a.getPoint3d();
            int i = 3 * a.getIndex();
            now[i] = p.x; now[i + 1] = p.y; now[i + 2] = p.z;
        }
        return now;
    }

    /** Addition (the reference has no gradient): dE/dx, [n][3], in the force field's energy unit per Angstrom. */
    public double[] calculateGradient(IAtomContainer mol) {
        calculateEnergy(mol);                                                 // binds, uploads
        try (Arena tmp = Arena.ofConfined()) {
            MemorySegment g = tmp.allocate(JAVA_DOUBLE, 3L * mol.getAtomCount());
            check((int) invoke(GRADIENT, handle, total, comps, g));
            return g.toArray(JAVA_DOUBLE);
        }
    }

    /**
     * The reference's minimizer loops run natively (jme_gd_minimize / jme_adam_minimize: GradientDescentMinimizer.java:
     * 49-105, AdamMinimizer.java:62-120 restated, no constraint): mode 0 = a full evaluation per trial like the Java
     * loops, 1 = jme_energy_delta per trial, 2 = the whole loop in one kernel (up to 8 192 atoms).  The container's
     * coordinates are updated; returns the energy after every step (index 0: the initial energy), as onStepConsumer
     * would have received them.  beta1 < 0 selects gradient descent, otherwise Adam with (beta1, beta2, epsilon).
     */
    public double[] minimizeNatively(IAtomContainer mol, int mode, int maxSteps, double stepSize, double threshold,
                                     double beta1, double beta2, double epsilon) {
        calculateEnergy(mol);                                                 // binds, marshals, uploads
        int n = mol.getAtomCount();
        try (Arena tmp = Arena.ofConfined()) {
            MemorySegment xyzInOut = tmp.allocate(JAVA_DOUBLE, 3L * n);
            MemorySegment.copy(coordinatesOf(mol), 0, xyzInOut, JAVA_DOUBLE, 0, 3 * n);
            MemorySegment calls = tmp.allocate(JAVA_LONG), steps = tmp.allocate(JAVA_INT);
            MemorySegment log = tmp.allocate(JAVA_DOUBLE, maxSteps + 2L);
            check((int) invoke(INCREMENTAL, handle, mode));
            if (beta1 < 0) check((int) invoke(GD, handle, maxSteps, stepSize, threshold, xyzInOut, calls, steps, log, maxSteps + 2));
            else check((int) invoke(ADAM, handle, maxSteps, stepSize, threshold, beta1, beta2, epsilon, xyzInOut, calls, steps, log, maxSteps + 2));
            double[] out = xyzInOut.toArray(JAVA_DOUBLE);
            for (IAtom a : mol.atoms()) {
                int i = 3 * a.getIndex();
                a.getPoint3d().set(out[i], out[i + 1], out[i + 2]);
            }
            shadow = out;                                                      // what the device holds now
            cachedOptions = null;
            return log.asSlice(0, 8L * (steps.get(JAVA_INT, 0) + 1)).toArray(JAVA_DOUBLE);
        }
    }

    private void uploadCoordinates(IAtomContainer mol) {
        int n = mol.getAtomCount(), changed = 0, last = -1;
        double[] now = coordinatesOf(mol);
        if (shadow != null) for (int k = 0; k < 3 * n; k++) if (now[k] != shadow[k]) { changed++; last = k; }
        if (shadow != null && changed == 0) return;
        if (shadow != null && changed == 1) {                                 // minimizer fast path: 8 bytes, not 24 n
            check((int) invoke(COORD1, handle, (long) (last / 3), last % 3, now[last]));
        } else {
            MemorySegment.copy(now, 0, xyz, JAVA_DOUBLE, 0, 3 * n);
            check((int) invoke(COORDS, handle, xyz, (long) n));
        }
        shadow = now;
    }

    /** Flatten what assignParameters left on the container (JmeSystemLayout.write) and create the device handle. */
    private void marshal(IAtomContainer mol, Object pairMap) {
        close0();
        arena.close();
        arena = Arena.ofShared();
        int n = mol.getAtomCount();
        try (Arena tmp = Arena.ofConfined()) {                                // jme_create copies everything it is given
            MemorySegment sys = JmeSystemLayout.write(tmp, mol);
            MemorySegment out = tmp.allocate(ADDRESS);
            int rc = (int) invoke(CREATE, sys, out);
            if (rc != 0) throw new RuntimeException(((MemorySegment) invoke(CREATEERR)).reinterpret(4096).getString(0));
            handle = out.get(ADDRESS, 0);
        }
        xyz = arena.allocate(JAVA_DOUBLE, 3L * n);
        total = arena.allocate(JAVA_DOUBLE);
        comps = arena.allocate(JAVA_DOUBLE, 7);
        delta = arena.allocate(JAVA_DOUBLE, 8);
        bound = mol; boundPairMap = pairMap; shadow = null; cachedOptions = null;
    }

    private void check(int rc) {
        if (rc != 0) throw new RuntimeException(((MemorySegment) invoke(LASTERR, handle)).reinterpret(4096).getString(0));
    }

    private static Object invoke(MethodHandle m, Object... a) {
        try { return m.invokeWithArguments(a); } catch (Throwable t) { throw new RuntimeException(t); }
    }

    private void close0() {
        if (!handle.equals(MemorySegment.NULL)) { invoke(DESTROY, handle); handle = MemorySegment.NULL; }
    }

    @Override
    public void close() { close0(); arena.close(); }
}
