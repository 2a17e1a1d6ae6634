package de.jstacs.sequenceScores.statisticalModels.trainable.hmm.models;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.ValueLayout;

import de.jstacs.algorithms.optimization.termination.AbstractTerminationCondition;
import de.jstacs.algorithms.optimization.termination.IterationCondition;
import de.jstacs.algorithms.optimization.termination.SmallDifferenceOfFunctionEvaluationsCondition;
import de.jstacs.data.DataSet;
import de.jstacs.data.WrongAlphabetException;
import de.jstacs.data.sequences.Sequence;
import de.jstacs.gpu.JstacsHmmNative;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.DifferentiableEmission;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.states.emissions.discrete.GpuEmissionExport;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.BaumWelchParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.MaxHMMTrainingParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.training.ViterbiParameterSet;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.BasicHigherOrderTransition;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.GpuTransitionExport;
import de.jstacs.sequenceScores.statisticalModels.trainable.hmm.transitions.elements.TransitionElement;
import de.jstacs.utils.IntList;
import de.jstacs.utils.Pair;
import de.jstacs.utils.Time;

/**
 * Drop-in replacement of {@link DifferentiableHigherOrderHMM} (the class <code>HMMFactory</code> returns for discrete
 * emissions, HMMFactory.java:111-119) whose data-set level methods run on a B200 through
 * <code>libjstacs_hmm_b200.so</code>:
 * <ul>
 * <li>{@link #train(DataSet, double[])} — the whole inner EM loop of HigherOrderHMM.java:733-789 (per start),</li>
 * <li>{@link #getLogScoreFor(DataSet, double[])} — HigherOrderHMM.java:927-941,</li>
 * <li>{@link #getViterbiPathsFor(DataSet)} — AbstractHMM.java:894-900,</li>
 * <li>{@link #decodePosteriorPathsFor(DataSet)} — getStatePosteriorMatrixFor + decodeStatePosterior (AbstractHMM.java:851-857, 1125-1139) fused.</li>
 * </ul>
 * Constructors, XML, cloning and every single-sequence method are inherited unchanged, so existing code that holds a
 * <code>TrainableStatisticalModel</code> or an <code>AbstractHMM</code> needs no edit beyond the <code>new</code>.
 * There is NO CPU fallback: if the library or a device is missing, or the model uses a feature outside the native
 * path (filters, reverse-strand states, states groups, position dependent transition elements, non-table emissions),
 * the call throws.
 */
public class GpuHigherOrderHMM extends DifferentiableHigherOrderHMM {

	public GpuHigherOrderHMM( MaxHMMTrainingParameterSet trainingParameterSet, String[] name, int[] emissionIdx, boolean[] forward,
			DifferentiableEmission[] emission, double ess, TransitionElement... te ) throws Exception {
		super( trainingParameterSet, name, emissionIdx, forward, emission, ess, te );
	}

	public GpuHigherOrderHMM( StringBuffer xml ) throws de.jstacs.io.NonParsableException {
		super( xml );
	}

	// ------------------------------------------------------------------------------------------------------------
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

	private static void put( MemorySegment d, String field, MemorySegment v ) {
		d.set( ValueLayout.ADDRESS, DESC.byteOffset( MemoryLayout.PathElement.groupElement( field ) ), v );
	}

	private static void put( MemorySegment d, String field, int v ) {
		d.set( ValueLayout.JAVA_INT, DESC.byteOffset( MemoryLayout.PathElement.groupElement( field ) ), v );
	}

	private int alphabet() {
		return (int) getAlphabetContainer().getAlphabetLengthAt( 0 );
	}

	/** Refuses what the native path does not cover instead of computing something else (SURVEY.md §8b "Errors"). */
	private void checkSupported() {
		if( !( transition instanceof BasicHigherOrderTransition ) ) {
			throw new UnsupportedOperationException( "only BasicHigherOrderTransition/HigherOrderTransition are on the native path" );
		}
		for( int i = 0; i < states.length; i++ ) {
			if( filter != null && filter[i] != null ) {
				throw new UnsupportedOperationException( "state filters are not on the native path" );
			}
			if( forward != null && !forward[i] ) {
				throw new UnsupportedOperationException( "reverse-strand states are not on the native path" );
			}
		}
		if( statesGroups != null ) {
			throw new UnsupportedOperationException( "allowed states groups are not on the native path" );
		}
		if( !getAlphabetContainer().isSimple() || !getAlphabetContainer().isDiscrete() ) {
			throw new UnsupportedOperationException( "only simple discrete alphabets are on the native path" );
		}
	}

	/** Flattens this model (Transition interface + emission tables) and creates the native handle. */
	private MemorySegment createModel( Arena arena ) throws Exception {
		checkSupported();
		GpuTransitionExport t = new GpuTransitionExport( (BasicHigherOrderTransition) transition );
		GpuEmissionExport e = new GpuEmissionExport( emission, alphabet() );
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
		put( d, "n_emis", emission.length );
		put( d, "state_emis", arena.allocateFrom( ValueLayout.JAVA_INT, emissionIdx ) );
		put( d, "emis_order", arena.allocateFrom( ValueLayout.JAVA_INT, e.emisOrder ) );
		put( d, "emis_cond_ptr", arena.allocateFrom( ValueLayout.JAVA_INT, e.emisCondPtr ) );
		put( d, "emis_param", arena.allocateFrom( ValueLayout.JAVA_DOUBLE, e.emisParam ) );
		put( d, "emis_lognorm", arena.allocateFrom( ValueLayout.JAVA_DOUBLE, e.emisLogNorm ) );
		put( d, "emis_hyper", arena.allocateFrom( ValueLayout.JAVA_DOUBLE, e.emisHyper ) );
		put( d, "state_silent", arena.allocateFrom( ValueLayout.JAVA_BYTE, silent ) );
		put( d, "state_final", arena.allocateFrom( ValueLayout.JAVA_BYTE, fin ) );
		MemorySegment out = arena.allocate( ValueLayout.ADDRESS );
		try {
			JstacsHmmNative.check( (int) JstacsHmmNative.jh_model_create.invokeExact( d, out ) );
		} catch( Exception ex ) {
			throw ex;
		} catch( Throwable th ) {
			throw new RuntimeException( th );
		}
		return out.get( ValueLayout.ADDRESS, 0 );
	}

	/**
	 * Packs a DataSet ONCE into uint8 codes + int64 offsets (off-heap, so the copy to the device needs no JNI pinning)
	 * — replaces the per-symbol virtual call <code>Sequence.discreteVal(int)</code> (Sequence.java:102) in the DP loops.
	 */
	private MemorySegment createDataSet( Arena arena, DataSet data, double[] weights ) throws Exception {
		if( !data.getAlphabetContainer().checkConsistency( getAlphabetContainer() ) ) {
			throw new WrongAlphabetException( "The AlphabetContainer of the data set and the model do not match." );
		}
		int n = data.getNumberOfElements();
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
		MemorySegment w = weights == null ? MemorySegment.NULL : arena.allocateFrom( ValueLayout.JAVA_DOUBLE, weights );
		MemorySegment out = arena.allocate( ValueLayout.ADDRESS );
		try {
			JstacsHmmNative.check( (int) JstacsHmmNative.jh_dataset_create.invokeExact( codes, offsets, (long) n, w, alphabet(), out ) );
		} catch( Exception ex ) {
			throw ex;
		} catch( Throwable th ) {
			throw new RuntimeException( th );
		}
		return out.get( ValueLayout.ADDRESS, 0 );
	}

	/** Parameters (and the statistics of the last E-step) computed on the device go back into the Java objects. */
	private void pullParameters( Arena arena, MemorySegment model ) throws Throwable {
		MemorySegment sz = arena.allocate( ValueLayout.JAVA_LONG, 4 );
		JstacsHmmNative.check( (int) JstacsHmmNative.jh_model_sizes.invokeExact( model, sz ) );
		long nChild = sz.getAtIndex( ValueLayout.JAVA_LONG, 0 ), nElem = sz.getAtIndex( ValueLayout.JAVA_LONG, 1 ),
				nEm = sz.getAtIndex( ValueLayout.JAVA_LONG, 2 ), nCond = sz.getAtIndex( ValueLayout.JAVA_LONG, 3 );
		MemorySegment tp = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( nChild, 1 ) ), tl = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( nElem, 1 ) ),
				ep = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( nEm, 1 ) ), el = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( nCond, 1 ) );
		JstacsHmmNative.check( (int) JstacsHmmNative.jh_model_get_params.invokeExact( model, tp, tl, ep, el ) );
		GpuTransitionExport.writeBack( (BasicHigherOrderTransition) transition, tp.toArray( ValueLayout.JAVA_DOUBLE ), tl.toArray( ValueLayout.JAVA_DOUBLE ), null );
		GpuEmissionExport e = new GpuEmissionExport( emission, alphabet() );
		GpuEmissionExport.writeBack( emission, e.emisCondPtr, alphabet(), ep.toArray( ValueLayout.JAVA_DOUBLE ), el.toArray( ValueLayout.JAVA_DOUBLE ), null );
		createStates(); // AbstractHMM.java:512 — states reference the (unchanged) emission objects
	}

	// ------------------------------------------------------------------------------------------------------------
	/**
	 * HigherOrderHMM.train (HigherOrderHMM.java:733-789): multi-start loop and best-of-starts stay here; the inner
	 * do-while (E-step, objective, termination test, M-step; :760-770) runs on the device.  The progress lines
	 * <code>it \t time \t value \t delta</code> (:764) are written after the fact from the per-iteration objectives.
	 */
	@Override
	public void train( DataSet data, double[] weights ) throws Exception {
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
		int numberOfStarts = trainingParameter.getNumberOfStarts();
		double best = Double.NEGATIVE_INFINITY;
		double[] bestParams = null;
		Time time = Time.getTimeInstance( sostream );
		try( Arena arena = Arena.ofConfined() ) {
			MemorySegment ds = createDataSet( arena, data, weights );
			try {
				for( int start = 0; start < numberOfStarts; start++ ) {
					sostream.writeln( "start " + start + " ============================" );
					if( !skipInit ) {
						initialize( data, weights );
					}
					MemorySegment model = createModel( arena ); // the (re-)initialised parameters travel with the descriptor
					try {
						int cap = kind == JstacsHmmNative.JH_TERM_ITERATIONS ? Math.max( maxIter, 1 ) : 100000;
						MemorySegment obj = arena.allocate( ValueLayout.JAVA_DOUBLE, cap ), nIt = arena.allocate( ValueLayout.JAVA_INT ),
								opts = arena.allocate( 24, 8 );
						opts.set( ValueLayout.JAVA_INT, 0, mode );
						opts.set( ValueLayout.JAVA_INT, 4, kind );
						opts.set( ValueLayout.JAVA_INT, 8, maxIter );
						opts.set( ValueLayout.JAVA_DOUBLE, 16, eps );
						time.reset();
						JstacsHmmNative.check( (int) JstacsHmmNative.jh_baum_welch.invokeExact( model, ds, opts, obj, cap, nIt ) );
						int n = Math.min( nIt.get( ValueLayout.JAVA_INT, 0 ), cap );
						double old = Double.NEGATIVE_INFINITY, cur = old;
						for( int it = 0; it < n; it++ ) {
							cur = obj.getAtIndex( ValueLayout.JAVA_DOUBLE, it );
							sostream.writeln( it + "\t" + time.getElapsedTime() + "\t" + cur + "\t" + ( cur - old ) );
							old = cur;
						}
						pullParameters( arena, model );
						if( cur > best ) { // :773-779
							best = cur;
							if( numberOfStarts > 1 ) {
								bestParams = getCurrentParameterValues();
							}
						}
					} finally {
						JstacsHmmNative.jh_model_destroy.invokeExact( model );
					}
				}
			} finally {
				JstacsHmmNative.jh_dataset_destroy.invokeExact( ds );
			}
		} catch( Exception e ) {
			throw e;
		} catch( Throwable t ) {
			throw new RuntimeException( t );
		}
		sostream.writeln( "best result: " + best );
		if( bestParams != null ) {
			setParameters( bestParams, 0 );
		}
	}

	/** HigherOrderHMM.getLogScoreFor(DataSet,double[]) (HigherOrderHMM.java:927-941), one launch for the whole data set. */
	@Override
	public void getLogScoreFor( DataSet data, double[] res ) throws Exception {
		int len = getLength(), l = data.getElementLength();
		if( len != 0 && l != len ) {
			throw new de.jstacs.data.WrongLengthException( "The length of the data set and the model do not match." );
		}
		try( Arena arena = Arena.ofConfined() ) {
			MemorySegment ds = createDataSet( arena, data, null ), model = createModel( arena );
			try {
				MemorySegment out = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( res.length, 1 ) );
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_loglik.invokeExact( model, ds, out ) );
				MemorySegment.copy( out, ValueLayout.JAVA_DOUBLE, 0, res, 0, data.getNumberOfElements() );
			} finally {
				JstacsHmmNative.jh_model_destroy.invokeExact( model );
				JstacsHmmNative.jh_dataset_destroy.invokeExact( ds );
			}
		} catch( Exception e ) {
			throw e;
		} catch( Throwable t ) {
			throw new RuntimeException( t );
		}
	}

	/**
	 * AbstractHMM.getViterbiPathsFor(DataSet) (AbstractHMM.java:894-900): bit-identical paths and scores.  Uses the
	 * general entry point <code>jh_viterbi_paths</code>, so models with silent states (paths of variable length,
	 * HigherOrderHMM.java:667-705) work as well.
	 */
	@Override
	@SuppressWarnings( "unchecked" )
	public Pair<IntList, Double>[] getViterbiPathsFor( DataSet data ) throws Exception {
		int n = data.getNumberOfElements(), nSilent = 0;
		for( int i = 0; i < states.length; i++ ) {
			nSilent += states[i].isSilent() ? 1 : 0;
		}
		Pair<IntList, Double>[] res = new Pair[n];
		try( Arena arena = Arena.ofConfined() ) {
			MemorySegment ds = createDataSet( arena, data, null ), model = createModel( arena );
			try {
				MemorySegment ptr = arena.allocate( ValueLayout.JAVA_LONG, n + 1L );
				long total = 0;
				for( int i = 0; i < n; i++ ) {
					ptr.setAtIndex( ValueLayout.JAVA_LONG, i, total );
					total += (long) data.getElementAt( i ).getLength() * ( 1 + nSilent ) + nSilent;
				}
				ptr.setAtIndex( ValueLayout.JAVA_LONG, n, total );
				MemorySegment paths = arena.allocate( ValueLayout.JAVA_INT, Math.max( total, 1 ) ), scores = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( n, 1 ) ),
						len = arena.allocate( ValueLayout.JAVA_LONG, Math.max( n, 1 ) );
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_viterbi_paths.invokeExact( model, ds, paths, ptr, len, scores ) );
				for( int i = 0; i < n; i++ ) {
					long p = ptr.getAtIndex( ValueLayout.JAVA_LONG, i ), L = len.getAtIndex( ValueLayout.JAVA_LONG, i );
					if( L < 0 ) {
						throw new IllegalArgumentException( "Impossible path" );
					}
					IntList path = new IntList( (int) L );
					for( long l = 0; l < L; l++ ) {
						path.add( paths.getAtIndex( ValueLayout.JAVA_INT, p + l ) );
					}
					res[i] = new Pair<IntList, Double>( path, scores.getAtIndex( ValueLayout.JAVA_DOUBLE, i ) );
				}
			} finally {
				JstacsHmmNative.jh_model_destroy.invokeExact( model );
				JstacsHmmNative.jh_dataset_destroy.invokeExact( ds );
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
		try( Arena arena = Arena.ofConfined() ) {
			MemorySegment ds = createDataSet( arena, data, null ), model = createModel( arena );
			try {
				MemorySegment out = arena.allocate( ValueLayout.JAVA_DOUBLE, Math.max( res.length, 1 ) );
				JstacsHmmNative.check( (int) JstacsHmmNative.jh_loglik_windows.invokeExact( model, ds, (long) res.length,
						arena.allocateFrom( ValueLayout.JAVA_LONG, seqIndex ), arena.allocateFrom( ValueLayout.JAVA_LONG, start ),
						arena.allocateFrom( ValueLayout.JAVA_LONG, end ), out ) );
				MemorySegment.copy( out, ValueLayout.JAVA_DOUBLE, 0, res, 0, res.length );
			} finally {
				JstacsHmmNative.jh_model_destroy.invokeExact( model );
				JstacsHmmNative.jh_dataset_destroy.invokeExact( ds );
			}
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
		try( Arena arena = Arena.ofConfined() ) {
			MemorySegment ds = createDataSet( arena, data, null ), model = createModel( arena );
			try {
				long total = 0;
				for( int i = 0; i < n; i++ ) {
					total += data.getElementAt( i ).getLength();
				}
				if( states.length <= 255 ) {
					// one byte per residue over PCIe (a quarter of the int32 traffic), widened here
					MemorySegment paths = arena.allocate( ValueLayout.JAVA_BYTE, Math.max( total, 1 ) );
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
					MemorySegment paths = arena.allocate( ValueLayout.JAVA_INT, Math.max( total, 1 ) );
					JstacsHmmNative.check( (int) JstacsHmmNative.jh_posterior_decode.invokeExact( model, ds, paths, MemorySegment.NULL ) );
					long p = 0;
					for( int i = 0; i < n; i++ ) {
						int L = data.getElementAt( i ).getLength();
						res[i] = new int[L];
						MemorySegment.copy( paths, ValueLayout.JAVA_INT, 4 * p, res[i], 0, L );
						p += L;
					}
				}
			} finally {
				JstacsHmmNative.jh_model_destroy.invokeExact( model );
				JstacsHmmNative.jh_dataset_destroy.invokeExact( ds );
			}
		} catch( Exception e ) {
			throw e;
		} catch( Throwable t ) {
			throw new RuntimeException( t );
		}
		return res;
	}
}
