package de.jstacs.sequenceScores.statisticalModels.trainable.hmm.models;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.ValueLayout;
import java.lang.ref.Cleaner;
import java.lang.ref.WeakReference;

import de.jstacs.algorithms.optimization.termination.AbstractTerminationCondition;
import de.jstacs.algorithms.optimization.termination.IterationCondition;
import de.jstacs.algorithms.optimization.termination.SmallDifferenceOfFunctionEvaluationsCondition;
import de.jstacs.data.AlphabetContainer;
import de.jstacs.data.DataSet;
import de.jstacs.data.WrongAlphabetException;
import de.jstacs.data.sequences.Sequence;
import de.jstacs.data.sequences.annotation.SequenceAnnotation;
import de.jstacs.gpu.JstacsHmmNative;
import de.jstacs.results.StorableResult;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.AbstractHMM;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.State;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.Emission;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.discrete.GpuEmissionExport;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.filter.Filter;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.BaumWelchParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.HMMTrainingParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.MaxHMMTrainingParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.ViterbiParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.BasicHigherOrderTransition;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.GpuTransitionExport;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.Transition;
import de.jstacs.utils.IntList;
import de.jstacs.utils.Pair;
import de.jstacs.utils.SafeOutputStream;
import de.jstacs.utils.Time;

/**
 * The part of the drop-in that talks to <code>libjstacs_hmm_b200.so</code>; shared by {@link GpuHigherOrderHMM}
 * (subclass of <code>DifferentiableHigherOrderHMM</code>, what <code>HMMFactory</code> returns) and
 * {@link GpuBasicHigherOrderHMM} (subclass of the plain <code>HigherOrderHMM</code> the cookbook builds,
 * supplementary/cookbook/Cookbook.java:416-419).  The model classes hand their <code>protected</code> state over
 * through {@link Host}; this class never reaches into them itself.
 * <p>
 * Handles are cached: the native model is created once per Java model (its structure never changes after the
 * constructor, AbstractHMM.java:245-309) and only the parameters are pushed before every call
 * (<code>jh_model_set_params</code>, a few KB); the packed data set is kept for the {@link DataSet} object it was built
 * from (Jstacs data sets are immutable), so "packed once, resident on the device" holds across calls, not only inside
 * <code>train</code>.  Call sequence per public method — replayed one to one by <code>tests/harness/shim_replay.c</code>:
 * <pre>
 *   first use          : jh_model_create(desc)
 *   new DataSet        : jh_dataset_destroy(old) ; jh_dataset_create(codes, offsets, n, weights, alphabet)
 *   same DataSet       : jh_dataset_set_weights(weights | NULL)
 *   every method       : jh_model_set_params(tp, tl, ep, el)
 *   train              : per start { [initialize ; jh_model_set_params] ; jh_baum_welch ; jh_last_objective ; jh_model_get_params }
 *   getLogScoreFor     : jh_loglik
 *   getViterbiPathsFor : jh_viterbi_u8 (no silent states, at most 255 states) | jh_viterbi_paths
 *   decodePosterior... : jh_posterior_decode_u8 | jh_posterior_decode
 *   getLogStatePost... : jh_posterior per sequence
 *   windows            : jh_loglik_windows
 *   close / GC         : jh_dataset_destroy ; jh_model_destroy
 * </pre>
 * There is NO CPU fallback: a missing library or device, or a model feature outside the native path, is an error.
 */
public final class GpuHmmEngine implements AutoCloseable {

	/** What the engine needs from the model object; implemented by the two shim classes, which see the protected members. */
	public interface Host {
		State[] gpuStates();                          // AbstractHMM.states            (AbstractHMM.java:75)
		Emission[] gpuEmission();                     // AbstractHMM.emission          (:101)
		int[] gpuEmissionIdx();                       // AbstractHMM.emissionIdx       (:90)
		boolean[] gpuForward();                       // AbstractHMM.forward           (:97, never null after the constructor :284-292)
		Filter[] gpuFilter();                         // AbstractHMM.filter            (:86, never null :259-263)
		Transition gpuTransition();                   // AbstractHMM.transition        (:108)
		boolean[] gpuFinalState();                    // AbstractHMM.finalState        (:135)
		int[][] gpuStatesGroups();                    // AbstractHMM.statesGroups      (:154; int[0][] for "no groups", :174-177)
		HMMTrainingParameterSet gpuTrainingParameter(); // AbstractHMM.trainingParameter (:125)
		SafeOutputStream gpuStream();                 // AbstractHMM.sostream          (:130)
		boolean gpuSkipInit();                        // HigherOrderHMM.skipInit       (HigherOrderHMM.java:128)
		String gpuAnnotationType();                   // HigherOrderHMM.type           (:136)
		void gpuInitialize( DataSet data, double[] weights ) throws Exception; // HigherOrderHMM.initialize (:867-869)
		void gpuCreateStates();                       // HigherOrderHMM.createStates   (:246-251)
		AlphabetContainer getAlphabetContainer();
		int getLength();
	}

	// jh_model_desc (include/jstacs_hmm.h): 4 x int32, 3 pointers, int32 (+pad), 8 pointers, int32 (+pad), 8 pointers
	private static final StructLayout DESC = MemoryLayout.structLayout(
			ValueLayout.JAVA_INT.withName( "abi_version" ), ValueLayout.JAVA_INT.withName( "n_states" ),
			ValueLayout.JAVA_INT.withName( "alphabet" ), ValueLayout.JAVA_INT.withName( "max_order" ),
			ValueLayout.ADDRESS.withName( "lc_ctx_ptr" ), ValueLayout.ADDRESS.withName( "ctx_elem" ), ValueLayout.ADDRESS.withName( "ctx_last_state" ),
			ValueLayout.JAVA_INT.withName( "n_elem" ), MemoryLayout.paddingLayout( 4 ),
			ValueLayout.ADDRESS.withName( "elem_child_ptr" ), ValueLayout.ADDRESS.withName( "child_state" ), ValueLayout.ADDRESS.withName( "child_desc" ),
			ValueLayout.ADDRESS.withName( "elem_ctx" ), ValueLayout.ADDRESS.withName( "trans_index" ), ValueLayout.ADDRESS.withName( "trans_param" ),
			ValueLayout.ADDRESS.withName( "trans_lognorm" ), ValueLayout.ADDRESS.withName( "trans_hyper" ),
			ValueLayout.JAVA_INT.withName( "n_emis" ), MemoryLayout.paddingLayout( 4 ),
			ValueLayout.ADDRESS.withName( "state_emis" ), ValueLayout.ADDRESS.withName( "emis_order" ), ValueLayout.ADDRESS.withName( "emis_cond_ptr" ),
			ValueLayout.ADDRESS.withName( "emis_param" ), ValueLayout.ADDRESS.withName( "emis_lognorm" ), ValueLayout.ADDRESS.withName( "emis_hyper" ),
			ValueLayout.ADDRESS.withName( "state_silent" ), ValueLayout.ADDRESS.withName( "state_final" ) );

	private static final Cleaner CLEANER = Cleaner.create();

	/** The native handles; released by {@link #close()} or, failing that, when the engine is garbage collected. */
	private static final class Handles implements Runnable {
		MemorySegment model = MemorySegment.NULL, dataset = MemorySegment.NULL;

		public void run() {
			try {
				if( !dataset.equals( MemorySegment.NULL ) ) {
					JstacsHmmNative.jh_dataset_destroy.invokeExact( dataset );
				}
				if( !model.equals( MemorySegment.NULL ) ) {
					JstacsHmmNative.jh_model_destroy.invokeExact( model );
				}
			} catch( Throwable t ) {
				// nothing sensible can be done in a cleaner
			}
			dataset = MemorySegment.NULL;
			model = MemorySegment.NULL;
		}
	}

	private final Host host;
	private final Handles handles = new Handles();
	private final Cleaner.Cleanable cleanable;
	private WeakReference<DataSet> cachedData = new WeakReference<DataSet>( null );
	private long cachedResidues;

	public GpuHmmEngine( Host host ) {
		this.host = host;
		this.cleanable = CLEANER.register( this, handles );
	}

	/** Releases the native model and data set; they are re-created on the next call. */
	@Override
	public void close() {
		handles.run();
		cachedData = new WeakReference<DataSet>( null );
	}

	private static void put( MemorySegment d, String field, MemorySegment v ) {
		d.set( ValueLayout.ADDRESS, DESC.byteOffset( MemoryLayout.PathElement.groupElement( field ) ), v );
	}

	private static void put( MemorySegment d, String field, int v ) {
		d.set( ValueLayout.JAVA_INT, DESC.byteOffset( MemoryLayout.PathElement.groupElement( field ) ), v );
	}

	private int alphabet() {
		return (int) host.getAlphabetContainer().getAlphabetLengthAt( 0 );
	}

	private int numberOfSilentStates() {
		int n = 0;
		for( State s : host.gpuStates() ) {
			n += s.isSilent() ? 1 : 0;
		}
		return n;
	}

	/** Refuses what the native path does not cover instead of computing something else (SURVEY.md §8b "Errors"). */
	private void checkSupported() {
		if( !( host.gpuTransition() instanceof BasicHigherOrderTransition ) ) {
			throw new UnsupportedOperationException( "only BasicHigherOrderTransition/HigherOrderTransition are on the native path" );
		}
		Filter[] filter = host.gpuFilter();
		boolean[] forward = host.gpuForward();
		for( int i = 0; i < host.gpuStates().length; i++ ) {
			if( filter != null && filter[i] != null ) {
				throw new UnsupportedOperationException( "state filters are not on the native path" );
			}
			if( forward != null && !forward[i] ) {
				throw new UnsupportedOperationException( "reverse-strand states are not on the native path" );
			}
		}
		// allowed states groups: "no groups" is an EMPTY array, never null (AbstractHMM.fillPreComputedContext,
		// AbstractHMM.java:173-177); declared groups go to the library in model() (jh_model_set_states_groups), which refuses
		// them for models with silent states or transition order 0
		AlphabetContainer con = host.getAlphabetContainer();
		if( !con.isSimple() || !con.isDiscrete() ) {
			throw new UnsupportedOperationException( "only simple discrete alphabets are on the native path" );
		}
	}

	/** Flattens the model (Transition tables + emission tables) and creates the native handle — once. */
	private MemorySegment model() throws Exception {
		if( !handles.model.equals( MemorySegment.NULL ) ) {
			return handles.model;
		}
		checkSupported();
		State[] states = host.gpuStates();
		boolean[] finalState = host.gpuFinalState();
		try( Arena arena = Arena.ofConfined() ) {
			GpuTransitionExport t = new GpuTransitionExport( (BasicHigherOrderTransition) host.gpuTransition() );
			GpuEmissionExport e = new GpuEmissionExport( host.gpuEmission(), alphabet() );
			byte[] silent = new byte[states.length], fin = new byte[states.length];
			for( int i = 0; i < states.length; i++ ) {
				silent[i] = (byte) ( states[i].isSilent() ? 1 : 0 ); // SimpleState.java:71-82
				fin[i] = (byte) ( finalState[i] ? 1 : 0 ); // AbstractHMM.java:1101-1113
			}
			MemorySegment d = arena.allocate( DESC );
			put( d, "abi_version", JstacsHmmNative.JH_ABI_VERSION );
			put( d, "n_states", states.length );
			put( d, "alphabet", alphabet() );
			put( d, "max_order", t.maxOrder );
			put( d, "lc_ctx_ptr", arena.allocateFrom( ValueLayout.JAVA_INT, t.lcCtxPtr ) );
			put( d, "ctx_elem", arena.allocateFrom( ValueLayout.JAVA_INT, t.ctxElem ) );
			put( d, "ctx_last_state", arena.allocateFrom( ValueLayout.JAVA_INT, t.ctxLastState ) );
			put( d, "n_elem", t.nElem );
			put( d, "elem_child_ptr", arena.allocateFrom( ValueLayout.JAVA_INT, t.elemChildPtr ) );
			put( d, "child_state", arena.allocateFrom( ValueLayout.JAVA_INT, t.childState ) );
			put( d, "child_desc", arena.allocateFrom( ValueLayout.JAVA_INT, t.childDesc ) );
			put( d, "elem_ctx", arena.allocateFrom( ValueLayout.JAVA_INT, t.elemCtx ) );
			put( d, "trans_index", arena.allocateFrom( ValueLayout.JAVA_INT, t.transIndex ) );
			put( d, "trans_param", arena.allocateFrom( ValueLayout.JAVA_DOUBLE, t.transParam ) );
			put( d, "trans_lognorm", arena.allocateFrom( ValueLayout.JAVA_DOUBLE, t.transLogNorm ) );
			put( d, "trans_hyper", arena.allocateFrom( ValueLayout.JAVA_DOUBLE, t.transHyper ) );
			put( d, "n_emis", host.gpuEmission().length );
			put( d, "state_emis", arena.allocateFrom( ValueLayout.JAVA_INT, host.gpuEmissionIdx() ) );
			put( d, "emis_order", arena.allocateFrom( ValueLayout.JAVA_INT, e.emisOrder ) );
			put( d, "emis_cond_ptr", arena.allocateFrom( ValueLayout.JAVA_INT, e.emisCondPtr ) );
			put( d, "emis_param", arena.allocateFrom( ValueLayout.JAVA_DOUBLE, e.emisParam ) );
			put( d, "emis_lognorm", arena.allocateFrom( ValueLayout.JAVA_DOUBLE, e.emisLogNorm ) );
			put( d, "emis_hyper", arena.allocateFrom( ValueLayout.JAVA_DOUBLE, e.emisHyper ) );
			put( d, "state_silent", arena.allocateFrom( ValueLayout.JAVA_BYTE, silent ) );
			put( d, "state_final", arena.allocateFrom( ValueLayout.JAVA_BYTE, fin ) );
			MemorySegment out = arena.allocate( ValueLayout.ADDRESS );
			JstacsHmmNative.check( (int) JstacsHmmNative.jh_model_create.invokeExact( d, out ) );
			handles.model = out.get( ValueLayout.ADDRESS, 0 );
			int[][] groups = host.gpuStatesGroups(); // the constructor argument int[][] statesGroups (AbstractHMM.java:154, 173-228)
			if( groups != null && groups.length > 0 ) {
				int[] ptr = new int[groups.length + 1];
				for( int g = 0; g < groups.length; g++ ) {
					ptr[g + 1] = ptr[g] + groups[g].length;
				}
				int[] members = new int[Math.max( ptr[groups.length], 1 )];
				for( int g = 0; g < groups.length; g++ ) {
					System.arraycopy( groups[g], 0, members, ptr[g], groups[g].length );
				}
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_model_set_states_groups.invokeExact( handles.model, groups.length,
						arena.allocateFrom( ValueLayout.JAVA_INT, ptr ), arena.allocateFrom( ValueLayout.JAVA_INT, members ) ) );
			}
		} catch( Exception ex ) {
			throw ex;
		} catch( Throwable th ) {
			throw new RuntimeException( th );
		}
		return handles.model;
	}

	/** The current Java-side parameters (after a CPU-side change, a random initialisation, setParameters, ...) go to the device. */
	private void pushParameters( Arena arena, MemorySegment model ) throws Throwable {
		GpuTransitionExport t = new GpuTransitionExport( (BasicHigherOrderTransition) host.gpuTransition() );
		GpuEmissionExport e = new GpuEmissionExport( host.gpuEmission(), alphabet() );
		JstacsHmmNative.check( (int) JstacsHmmNative.jh_model_set_params.invokeExact( model,
				arena.allocateFrom( ValueLayout.JAVA_DOUBLE, t.transParam ), arena.allocateFrom( ValueLayout.JAVA_DOUBLE, t.transLogNorm ),
				arena.allocateFrom( ValueLayout.JAVA_DOUBLE, e.emisParam ), arena.allocateFrom( ValueLayout.JAVA_DOUBLE, e.emisLogNorm ) ) );
	}

	/** Parameters computed on the device, in the layout of jh_model_desc: { trans_param, trans_lognorm, emis_param, emis_lognorm }. */
	private double[][] fetchParameters( Arena arena, MemorySegment model ) throws Throwable {
		MemorySegment sz = arena.allocate( ValueLayout.JAVA_LONG, 4 );
		JstacsHmmNative.check( (int) JstacsHmmNative.jh_model_sizes.invokeExact( model, sz ) );
		long nChild = sz.getAtIndex( ValueLayout.JAVA_LONG, 0 ), nElem = sz.getAtIndex( ValueLayout.JAVA_LONG, 1 ),
				nEm = sz.getAtIndex( ValueLayout.JAVA_LONG, 2 ), nCond = sz.getAtIndex( ValueLayout.JAVA_LONG, 3 );
		MemorySegment tp = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( nChild, 1 ) ), tl = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( nElem, 1 ) ),
				ep = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( nEm, 1 ) ), el = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( nCond, 1 ) );
		JstacsHmmNative.check( (int) JstacsHmmNative.jh_model_get_params.invokeExact( model, tp, tl, ep, el ) );
		return new double[][]{ tp.toArray( ValueLayout.JAVA_DOUBLE ), tl.toArray( ValueLayout.JAVA_DOUBLE ), ep.toArray( ValueLayout.JAVA_DOUBLE ), el.toArray( ValueLayout.JAVA_DOUBLE ) };
	}

	/** Device parameters go back into the Java objects, so toXML(), clone() and every inherited CPU method keep working. */
	private void writeBack( double[][] p ) {
		GpuTransitionExport.writeBack( (BasicHigherOrderTransition) host.gpuTransition(), p[0], p[1], null );
		GpuEmissionExport e = new GpuEmissionExport( host.gpuEmission(), alphabet() );
		GpuEmissionExport.writeBack( host.gpuEmission(), e.emisCondPtr, alphabet(), p[2], p[3], null );
		host.gpuCreateStates(); // as HigherOrderHMM.train does after replacing the emissions (HigherOrderHMM.java:785)
	}

	/**
	 * Packs a DataSet ONCE into uint8 codes + int64 offsets (off-heap, so the copy to the device needs no JNI pinning; the
	 * library stages it through its own pinned buffers) — replaces the per-symbol virtual call
	 * <code>Sequence.discreteVal(int)</code> (Sequence.java:102) in the DP loops.  The handle is kept for this DataSet.
	 */
	private MemorySegment dataset( DataSet data, double[] weights ) throws Exception {
		if( !data.getAlphabetContainer().checkConsistency( host.getAlphabetContainer() ) ) {
			throw new WrongAlphabetException( "The AlphabetContainer of the data set and the model do not match." );
		}
		int n = data.getNumberOfElements();
		if( weights != null && weights.length != n ) {
			throw new IllegalArgumentException( "The weights do not match the data set." );
		}
		try( Arena arena = Arena.ofConfined() ) {
			MemorySegment w = weights == null ? MemorySegment.NULL : arena.allocateFrom( ValueLayout.JAVA_DOUBLE, weights );
			if( cachedData.get() == data && !handles.dataset.equals( MemorySegment.NULL ) ) {
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_dataset_set_weights.invokeExact( handles.dataset, w ) );
				// per-position groups are attached by train() only; every other call computes the unconstrained quantity,
				// like the reference methods without an allowedStatesGroup argument
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_dataset_set_groups.invokeExact( handles.dataset, MemorySegment.NULL ) );
				return handles.dataset;
			}
			if( !handles.dataset.equals( MemorySegment.NULL ) ) {
				JstacsHmmNative.jh_dataset_destroy.invokeExact( handles.dataset );
				handles.dataset = MemorySegment.NULL;
			}
			long total = 0;
			for( int i = 0; i < n; i++ ) {
				total += data.getElementAt( i ).getLength();
			}
			MemorySegment codes = arena.allocate( Math.max( total, 1 ), 16 );
			MemorySegment offsets = arena.allocate( ValueLayout.JAVA_LONG, n + 1L );
			long p = 0;
			for( int i = 0; i < n; i++ ) {
				Sequence s = data.getElementAt( i );
				offsets.setAtIndex( ValueLayout.JAVA_LONG, i, p );
				for( int l = 0, len = s.getLength(); l < len; l++, p++ ) {
					codes.set( ValueLayout.JAVA_BYTE, p, (byte) s.discreteVal( l ) );
				}
			}
			offsets.setAtIndex( ValueLayout.JAVA_LONG, n, p );
			MemorySegment out = arena.allocate( ValueLayout.ADDRESS );
			JstacsHmmNative.check( (int) JstacsHmmNative.jh_dataset_create.invokeExact( codes, offsets, (long) n, w, alphabet(), out ) );
			handles.dataset = out.get( ValueLayout.ADDRESS, 0 );
			cachedData = new WeakReference<DataSet>( data );
			cachedResidues = total;
			return handles.dataset;
		} catch( Exception ex ) {
			throw ex;
		} catch( Throwable th ) {
			throw new RuntimeException( th );
		}
	}

	/**
	 * HigherOrderHMM.doOneStep (HigherOrderHMM.java:810-820): with a sequence annotation type, every sequence that carries an
	 * annotation of that type is trained under its AllowedStatesGroups (one group index per position, -1 = unconstrained);
	 * sequences without the annotation are unconstrained.  The per-position indices go to the data set handle.
	 */
	private void attachGroups( Arena arena, MemorySegment ds, DataSet data ) throws Throwable {
		String type = host.gpuAnnotationType();
		if( type == null ) {
			return; // dataset() has detached whatever an earlier call attached
		}
		MemorySegment g = arena.allocate( ValueLayout.JAVA_INT, Math.max( cachedResidues, 1 ) );
		long p = 0;
		boolean any = false;
		for( int n = 0; n < data.getNumberOfElements(); n++ ) {
			Sequence seq = data.getElementAt( n );
			int[] groups = null;
			SequenceAnnotation sa = seq.getSequenceAnnotationByType( type, 0 );
			if( sa != null ) {
				StorableResult res = (StorableResult) sa.getResultAt( 0 );
				groups = ( (AbstractHMM.AllowedStatesGroups) res.getResultInstance() ).groups;
				if( groups.length != seq.getLength() ) {
					throw new IllegalArgumentException( "AllowedStatesGroups of sequence " + n + " does not match its length." );
				}
				any = true;
			}
			for( int l = 0, len = seq.getLength(); l < len; l++, p++ ) {
				g.setAtIndex( ValueLayout.JAVA_INT, p, groups == null ? -1 : groups[l] );
			}
		}
		if( any ) {
			JstacsHmmNative.check( (int) JstacsHmmNative.jh_dataset_set_groups.invokeExact( ds, g ) );
		}
	}

	// ------------------------------------------------------------------------------------------------------------
	/**
	 * HigherOrderHMM.train (HigherOrderHMM.java:733-789): multi-start loop and best-of-starts stay here; the inner
	 * do-while (E-step, objective, termination test, M-step; :760-770) runs on the device.  The progress lines
	 * <code>it \t time \t value \t delta</code> (:764) are written after the fact from the per-iteration objectives.
	 */
	public void train( DataSet data, double[] weights ) throws Exception {
		HMMTrainingParameterSet trainingParameter = host.gpuTrainingParameter();
		if( !( trainingParameter instanceof MaxHMMTrainingParameterSet ) ) {
			throw new IllegalArgumentException( "This kind of training is currently not supported." );
		}
		int mode;
		if( trainingParameter instanceof ViterbiParameterSet ) {
			mode = JstacsHmmNative.JH_MODE_VITERBI;
		} else if( trainingParameter instanceof BaumWelchParameterSet ) {
			mode = JstacsHmmNative.JH_MODE_BAUM_WELCH;
		} else {
			throw new IllegalArgumentException( "Training mode not available." ); // HigherOrderHMM.java:805
		}
		AbstractTerminationCondition tc = ( (MaxHMMTrainingParameterSet) trainingParameter ).getTerminationCondition();
		int kind, maxIter = 0;
		double eps = 0;
		if( tc instanceof IterationCondition ) {
			kind = JstacsHmmNative.JH_TERM_ITERATIONS;
			maxIter = (Integer) tc.getCurrentParameterSet().getParameterAt( 0 ).getValue();
		} else if( tc instanceof SmallDifferenceOfFunctionEvaluationsCondition ) {
			kind = JstacsHmmNative.JH_TERM_SMALL_DIFFERENCE;
			eps = (Double) tc.getCurrentParameterSet().getParameterAt( 0 ).getValue();
		} else {
			throw new UnsupportedOperationException( "termination condition " + tc.getClass().getSimpleName() + " is not on the native path" );
		}
		SafeOutputStream sostream = host.gpuStream();
		int numberOfStarts = trainingParameter.getNumberOfStarts();
		double best = Double.NEGATIVE_INFINITY;
		double[][] bestParams = null;
		Time time = Time.getTimeInstance( sostream );
		MemorySegment model = model(), ds = dataset( data, weights );
		try( Arena arena = Arena.ofConfined() ) {
			attachGroups( arena, ds, data );
			int cap = kind == JstacsHmmNative.JH_TERM_ITERATIONS ? Math.max( maxIter, 1 ) : 100000;
			MemorySegment obj = arena.allocate( ValueLayout.JAVA_DOUBLE, cap ), nIt = arena.allocate( ValueLayout.JAVA_INT ),
					last = arena.allocate( ValueLayout.JAVA_DOUBLE ), opts = arena.allocate( 24, 8 );
			opts.set( ValueLayout.JAVA_INT, 0, mode );
			opts.set( ValueLayout.JAVA_INT, 4, kind );
			opts.set( ValueLayout.JAVA_INT, 8, maxIter );
			opts.set( ValueLayout.JAVA_DOUBLE, 16, eps );
			for( int start = 0; start < numberOfStarts; start++ ) {
				sostream.writeln( "start " + start + " ============================" );
				if( !host.gpuSkipInit() ) {
					host.gpuInitialize( data, weights );
				}
				pushParameters( arena, model );
				time.reset();
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_baum_welch.invokeExact( model, ds, opts, obj, cap, nIt ) );
				int n = Math.min( nIt.get( ValueLayout.JAVA_INT, 0 ), cap );
				double old = Double.NEGATIVE_INFINITY, cur;
				for( int it = 0; it < n; it++ ) {
					cur = obj.getAtIndex( ValueLayout.JAVA_DOUBLE, it );
					sostream.writeln( it + "\t" + time.getElapsedTime() + "\t" + cur + "\t" + ( cur - old ) );
					old = cur;
				}
				// the objective of the last iteration, whatever the capacity of the history buffer
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_last_objective.invokeExact( model, last ) );
				double newValue = last.get( ValueLayout.JAVA_DOUBLE, 0 );
				double[][] params = fetchParameters( arena, model );
				writeBack( params );
				if( newValue > best ) { // :773-779
					best = newValue;
					if( numberOfStarts > 1 ) {
						bestParams = params;
					}
				}
			}
		} catch( Exception e ) {
			throw e;
		} catch( Throwable t ) {
			throw new RuntimeException( t );
		}
		sostream.writeln( "best result: " + best );
		if( bestParams != null ) {
			writeBack( bestParams );
		}
	}

	/** HigherOrderHMM.getLogScoreFor(DataSet,double[]) (HigherOrderHMM.java:927-941), one launch for the whole data set. */
	public void getLogScoreFor( DataSet data, double[] res ) throws Exception {
		int len = host.getLength(), l = data.getElementLength();
		if( len != 0 && l != len ) {
			throw new de.jstacs.data.WrongLengthException( "The length of the data set and the model do not match." );
		}
		MemorySegment model = model(), ds = dataset( data, null );
		try( Arena arena = Arena.ofConfined() ) {
			pushParameters( arena, model );
			MemorySegment out = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( res.length, 1 ) );
			JstacsHmmNative.check( (int) JstacsHmmNative.jh_loglik.invokeExact( model, ds, out ) );
			MemorySegment.copy( out, ValueLayout.JAVA_DOUBLE, 0, res, 0, data.getNumberOfElements() );
		} catch( Exception e ) {
			throw e;
		} catch( Throwable t ) {
			throw new RuntimeException( t );
		}
	}

	/**
	 * AbstractHMM.getViterbiPathsFor(DataSet) (AbstractHMM.java:894-900): bit-identical paths and scores.  Models without
	 * silent states emit one state per residue and take <code>jh_viterbi_u8</code> (one byte per residue over PCIe);
	 * models with silent states (paths of variable length, HigherOrderHMM.java:667-705) take <code>jh_viterbi_paths</code>.
	 * Both run the same kernels on the device for models without silent states.
	 */
	@SuppressWarnings( "unchecked" )
	public Pair<IntList, Double>[] getViterbiPathsFor( DataSet data ) throws Exception {
		int n = data.getNumberOfElements(), nSilent = numberOfSilentStates();
		Pair<IntList, Double>[] res = new Pair[n];
		MemorySegment model = model(), ds = dataset( data, null );
		try( Arena arena = Arena.ofConfined() ) {
			pushParameters( arena, model );
			MemorySegment scores = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( n, 1 ) );
			if( nSilent == 0 && host.gpuStates().length <= 255 ) {
				MemorySegment paths = arena.allocate( Math.max( cachedResidues, 1 ), 16 );
				// JH_ERR_IMPOSSIBLE -> IllegalArgumentException("Impossible path"), like HigherOrderHMM.java:659
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_viterbi_u8.invokeExact( model, ds, paths, scores ) );
				long p = 0;
				for( int i = 0; i < n; i++ ) {
					int L = data.getElementAt( i ).getLength();
					IntList path = new IntList( Math.max( L, 1 ) );
					for( int l = 0; l < L; l++ ) {
						path.add( paths.get( ValueLayout.JAVA_BYTE, p + l ) & 0xFF );
					}
					p += L;
					res[i] = new Pair<IntList, Double>( path, scores.getAtIndex( ValueLayout.JAVA_DOUBLE, i ) );
				}
			} else {
				MemorySegment ptr = arena.allocate( ValueLayout.JAVA_LONG, n + 1L );
				long total = 0;
				for( int i = 0; i < n; i++ ) {
					ptr.setAtIndex( ValueLayout.JAVA_LONG, i, total );
					total += (long) data.getElementAt( i ).getLength() * ( 1 + nSilent ) + nSilent;
				}
				ptr.setAtIndex( ValueLayout.JAVA_LONG, n, total );
				MemorySegment paths = arena.allocate( ValueLayout.JAVA_INT, Math.max( total, 1 ) ), len = arena.allocate( ValueLayout.JAVA_LONG, Math.max( n, 1 ) );
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_viterbi_paths.invokeExact( model, ds, paths, ptr, len, scores ) );
				for( int i = 0; i < n; i++ ) {
					long p = ptr.getAtIndex( ValueLayout.JAVA_LONG, i ), L = len.getAtIndex( ValueLayout.JAVA_LONG, i );
					if( L < 0 ) {
						throw new IllegalArgumentException( "Impossible path" );
					}
					IntList path = new IntList( (int) Math.max( L, 1 ) );
					for( long l = 0; l < L; l++ ) {
						path.add( paths.getAtIndex( ValueLayout.JAVA_INT, p + l ) );
					}
					res[i] = new Pair<IntList, Double>( path, scores.getAtIndex( ValueLayout.JAVA_DOUBLE, i ) );
				}
			}
		} catch( Exception e ) {
			throw e;
		} catch( Throwable t ) {
			throw new RuntimeException( t );
		}
		return res;
	}

	/**
	 * AbstractHMM.getLogStatePosteriorMatrixFor(DataSet) (AbstractHMM.java:831-838): <code>double[N][S][L_n]</code>, one
	 * <code>jh_posterior</code> per sequence (the matrices are the product here, S x L doubles per sequence cross PCIe).
	 */
	public double[][][] getLogStatePosteriorMatrixFor( DataSet data ) throws Exception {
		int n = data.getNumberOfElements(), S = host.gpuStates().length;
		double[][][] res = new double[n][S][];
		MemorySegment model = model(), ds = dataset( data, null );
		try( Arena arena = Arena.ofConfined() ) {
			pushParameters( arena, model );
			for( int i = 0; i < n; i++ ) {
				int L = data.getElementAt( i ).getLength();
				try( Arena a2 = Arena.ofConfined() ) {
					MemorySegment out = a2.allocate( ValueLayout.JAVA_DOUBLE, Math.max( (long) S * L, 1 ) );
					JstacsHmmNative.check( (int) JstacsHmmNative.jh_posterior.invokeExact( model, ds, (long) i, out ) );
					for( int s = 0; s < S; s++ ) {
						res[i][s] = new double[L];
						MemorySegment.copy( out, ValueLayout.JAVA_DOUBLE, 8L * s * L, res[i][s], 0, L );
					}
				}
			}
		} catch( Exception e ) {
			throw e;
		} catch( Throwable t ) {
			throw new RuntimeException( t );
		}
		return res;
	}

	/**
	 * <code>getLogProbFor(data.getElementAt(seqIndex[i]), start[i], end[i])</code> (AbstractHMM.java:985-998) for many
	 * windows in one call — the sliding-window scan of projects/xanthogenomes/NHMMer.java:225-244.
	 */
	public double[] getLogProbForWindows( DataSet data, long[] seqIndex, long[] start, long[] end ) throws Exception {
		double[] res = new double[seqIndex.length];
		MemorySegment model = model(), ds = dataset( data, null );
		try( Arena arena = Arena.ofConfined() ) {
			pushParameters( arena, model );
			MemorySegment out = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( res.length, 1 ) );
			JstacsHmmNative.check( (int) JstacsHmmNative.jh_loglik_windows.invokeExact( model, ds, (long) res.length,
					arena.allocateFrom( ValueLayout.JAVA_LONG, seqIndex ), arena.allocateFrom( ValueLayout.JAVA_LONG, start ),
					arena.allocateFrom( ValueLayout.JAVA_LONG, end ), out ) );
			MemorySegment.copy( out, ValueLayout.JAVA_DOUBLE, 0, res, 0, res.length );
		} catch( Exception e ) {
			throw e;
		} catch( Throwable t ) {
			throw new RuntimeException( t );
		}
		return res;
	}

	/**
	 * <code>decodeStatePosterior(getStatePosteriorMatrixFor(seq))</code> for every sequence (AbstractHMM.java:810-857,
	 * 1125-1139) without materialising the S x L matrices on the host.
	 */
	public int[][] decodePosteriorPathsFor( DataSet data ) throws Exception {
		int n = data.getNumberOfElements();
		int[][] res = new int[n][];
		MemorySegment model = model(), ds = dataset( data, null );
		try( Arena arena = Arena.ofConfined() ) {
			pushParameters( arena, model );
			if( host.gpuStates().length <= 255 ) {
				// one byte per residue over PCIe (a quarter of the int32 traffic), widened here
				MemorySegment paths = arena.allocate( Math.max( cachedResidues, 1 ), 16 );
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_posterior_decode_u8.invokeExact( model, ds, paths, MemorySegment.NULL ) );
				long p = 0;
				for( int i = 0; i < n; i++ ) {
					int L = data.getElementAt( i ).getLength();
					res[i] = new int[L];
					for( int l = 0; l < L; l++ ) {
						res[i][l] = paths.get( ValueLayout.JAVA_BYTE, p + l ) & 0xFF;
					}
					p += L;
				}
			} else {
				MemorySegment paths = arena.allocate( ValueLayout.JAVA_INT, Math.max( cachedResidues, 1 ) );
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_posterior_decode.invokeExact( model, ds, paths, MemorySegment.NULL ) );
				long p = 0;
				for( int i = 0; i < n; i++ ) {
					int L = data.getElementAt( i ).getLength();
					res[i] = new int[L];
					MemorySegment.copy( paths, ValueLayout.JAVA_INT, 4 * p, res[i], 0, L );
					p += L;
				}
			}
		} catch( Exception e ) {
			throw e;
		} catch( Throwable t ) {
			throw new RuntimeException( t );
		}
		return res;
	}
}
